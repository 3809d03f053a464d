// K7/K8/K9: STFT family -- placeholder entry points (fail loudly) until the kernels land.
#include "common.cuh"
using namespace b200;
extern "C" {
#define NOT_YET(name) set_error(name ": kernel not built yet"); return B200VOC_ERR_UNSUPPORTED
int b200voc_stft_mag(const float*, int, int, int, int, const float*, float*, void*) { NOT_YET("stft_mag"); }
int b200voc_stft_complex(const float*, int, int, int, int, float*, void*) { NOT_YET("stft_complex"); }
int b200voc_stft_logmel(const float*, int, int, int, int, int, int, int, float*, void*) { NOT_YET("stft_logmel"); }
int b200voc_istft(const float*, int, int, int, int, int, float*, void*) { NOT_YET("istft"); }
int b200voc_stft_l1(const float*, const float*, int, int, int, int, const float*, double*, void*) { NOT_YET("stft_l1"); }
}
