/* b200voc_dev.h -- experiment / debug entry points of the DEVELOPMENT build (libb200voc_dev.so, `make dev`:
 * -DB200VOC_DEV, which also compiles the clock64 trace hooks and knock-out switches into the kernels).  None of
 * these is exported by the product library libb200voc.so; tests/exp_*.py, tests/trace_*.py, tests/knock_*.py and
 * tests/gpu_probe.py select the dev build with B200VOC_LIB=dev. */
#ifndef B200VOC_DEV_H_
#define B200VOC_DEV_H_
#include "b200voc.h"
#ifdef __cplusplus
extern "C" {
#endif

/* experiment: UMMA descriptors whose start address is offset by whole 128B rows. */
int b200voc_exp_rowshift(const void* a16_144x64, const void* b16_64x64, float* out_2x16x128x64, void* stream);
/* debug: when non-NULL, the narrow-stage residual-block kernel records a clock64 timeline of CTA 0
 * ([5 roles][64 tiles][4] int64) into dev_buf; NULL switches it off. */
int b200voc_debug_set_trace(int64_t* dev_buf);
/* experiment: one cta_group::2 MMA group per CTA pair, D[256*pairs,128] = A[256*pairs,64] B[128,64]^T (fp16 in,
 * fp32 out); cycles[pairs] (optional) = issue -> completion seen by the leader. */
int b200voc_exp_cta2(const void* a16, const void* b16, int pairs, float* out, int64_t* cycles, void* stream);
/* experiment: cycles for iters x 4 tcgen05.mma (M=128, N=n, K=16, operands in shared memory). */
int b200voc_exp_mma_rate(int n, int iters, int blocks, int64_t* out_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VOC_DEV_H_ */
