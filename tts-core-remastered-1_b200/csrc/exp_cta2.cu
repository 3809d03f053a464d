// Experiment (not on the product path): a CTA pair (cluster of 2) cooperating on one tcgen05.mma
// with cta_group::2 -- M = 256 (128 rows per CTA), N = 128, each CTA holding HALF of the B operand.
// Verifies the PTX forms a 2-CTA version of the wide-stage residual block would need: cluster
// launch, cta_group::2 TMEM allocation, the M=256 instruction descriptor, remote mbarrier arrive
// (mapa), tcgen05.commit multicast to both CTAs.  D[256 x 128] = A[256 x 64] * B[128 x 64]^T.
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
exp_cta2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
                long long* cycles, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                      // 128 rows x 128 B (this CTA's half of M)
  uint8_t* sB = smem + 16384;              // 64 rows x 128 B  (this CTA's half of N)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint64_t* loaded = bars;                 // local TMA bytes landed
  uint64_t* peer_ready = bars + 1;         // (leader) the peer's operands have landed
  uint64_t* done = bars + 2;               // MMAs complete (commit multicast arrives here in BOTH CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    mbar_init(loaded, 1);
    mbar_init(peer_ready, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_2cta(tmem_slot, 128);
  tc_fence_before();
  cluster_sync_all();                      // barrier inits + TMEM allocation visible to the peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long t0 = 0;
  if (threadIdx.x == 0 && variant == 1) {
    // variant 1: both CTAs' TMA loads signal the leader's barrier (cta_group::2 TMA form)
    if (rank == 0) mbar_expect_tx(loaded, 2 * (16384 + 8192));
    tma_load_2d_2cta(sA, &tmA, loaded, 0, pair * 256 + rank * 128);
    tma_load_2d_2cta(sB, &tmB, loaded, 0, rank * 64);
    if (rank == 0) {
      mbar_wait(loaded, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16_m(0, 256, 128);
      const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sA));
      const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sB));
      t0 = clock64();
      for (int k = 0; k < 4; ++k) umma_f16_2cta(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
      umma_commit_2cta(done, 0x3);
    }
  } else if (threadIdx.x == 0) {
    mbar_expect_tx(loaded, 16384 + 8192);
    tma_load_2d(sA, &tmA, loaded, 0, pair * 256 + rank * 128);
    tma_load_2d(sB, &tmB, loaded, 0, rank * 64);
    mbar_wait(loaded, 0);
    if (rank == 1) {
      mbar_arrive_remote(mapa_u32(peer_ready, 0));       // tell the leader our halves are in place
    } else {
      mbar_wait(peer_ready, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16_m(0, 256, 128);
      const uint64_t a_desc = make_kmajor_desc<128>(smem_u32(sA));
      const uint64_t b_desc = make_kmajor_desc<128>(smem_u32(sB));
      t0 = clock64();
      for (int k = 0; k < 4; ++k) umma_f16_2cta(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
      umma_commit_2cta(done, 0x3);
    }
  }
  mbar_wait(done, 0);
  if (threadIdx.x == 0 && rank == 0 && cycles) cycles[pair] = clock64() - t0;
  tc_fence_after();
  float* o = out + ((long long)pair * 256 + rank * 128 + warp * 32 + lane) * 128;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) o[c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  cluster_sync_all();                      // both CTAs are done with TMEM before it is released
  if (warp == 0) tmem_dealloc_2cta(tmem_base, 128);
}

int exp_cta2_launch(const void* a16, const void* b16, int pairs, float* out, long long* cycles, cudaStream_t stream) {
  const int variant = pairs >= 1000 ? 1 : 0;
  pairs %= 1000;
  CUtensorMap tmA, tmB;
  B200_TRY(make_tmap_2d(&tmA, a16, 64, 256ull * pairs, 128, 64, 128, 128));
  B200_TRY(make_tmap_2d(&tmB, b16, 64, 128, 128, 64, 64, 128));
  const int SMEM = 16384 + 8192 + 256 + 1024;
  exp_cta2_kernel<<<2 * pairs, 128, SMEM, stream>>>(tmA, tmB, out, cycles, variant);
  B200_CUDA(cudaGetLastError());
  return B200VOC_OK;
}

}  // namespace b200
