// K6: SelfAttention (generator.py:43-44,91-92; body = repair R3).  Placeholder until the
// flash-style tcgen05 kernel lands: the entry point fails loudly, there is no fallback.
#include "common.cuh"

namespace b200 {

long long attention_scratch_elems(int N, int L, int C) { return 4ll * N * L * C; }

int attention_launch(const void*, const void*, const float*, const void*, const float*, int, int, int, int, int, void*,
                     void*, void*, void*, void*, cudaStream_t) {
  set_error("SelfAttention CUDA kernel (K6) is not built yet: construct the Generator with use_attention=False");
  return B200VOC_ERR_UNSUPPORTED;
}

}  // namespace b200
