// Probe for the MN-major B operand of tcgen05.mma (needed by a tcgen05 P·V attention GEMM, where V sits in the
// paged cache as [key][d], i.e. with the contraction dimension along ROWS). One CTA computes
//     D[128, 128] = A[128, 64] · Bt[64, 128]          (A K-major / SW128 as in gemm_bf16.cu; Bt row-major)
// with the B shared-memory descriptor fields (LBO, SBO), the per-UMMA_K start-address advance and the stride
// between the two 64-wide N chunks given at RUN TIME, so that one GPU call can sweep the candidate encodings
// (benchmarks/umma_mn_sweep.py) and report which one reproduces torch. Not used by the product path.
#include "../common/host_utils.h"
#include "../common/ptx.cuh"

namespace b200 {

struct MnProbeParams {
  float* D;                 // [128, 128] fp32 out
  uint32_t lbo16, sbo16;    // descriptor fields, in 16-byte units
  uint32_t k_adv16;         // start-address advance per UMMA_K (16 rows of K), 16-byte units
  uint32_t n_chunk_bytes;   // byte offset of the second 64-wide N chunk inside the B tile in smem
  uint32_t b_major;         // instruction-descriptor bit 16
  uint32_t split_n;         // 1: issue two N=64 MMAs (one per chunk) instead of one N=128 MMA
};

__global__ void __launch_bounds__(128, 1)
umma_mn_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const MnProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                  // A: 128 rows x 64 cols bf16 = 16 KB (K-major SW128)
  uint8_t* sb = smem + 16384;          // B: two boxes of [64 K rows][64 N cols] bf16 = 2 x 8 KB (MN-major SW128)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  uint64_t* done = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<1>(tmem_ptr, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 32768);
    tma_load_2d(sa, &tmap_a, bar, 0, 0, kEvictNormal);             // A[0:128, 0:64]
    tma_load_2d(sb, &tmap_b, bar, 0, 0, kEvictNormal);             // Bt[0:64, 0:64]
    tma_load_2d(sb + 8192, &tmap_b, bar, 64, 0, kEvictNormal);     // Bt[0:64, 64:128]
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint64_t da = make_sw128_kmajor_desc(smem_u32(sa));
    auto make_b = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
      d |= static_cast<uint64_t>(p.lbo16 & 0x3FFF) << 16;
      d |= static_cast<uint64_t>(p.sbo16 & 0x3FFF) << 32;
      d |= static_cast<uint64_t>(1) << 46;
      d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
      return d;
    };
    const uint32_t n = p.split_n ? 64u : 128u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (p.b_major << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {            // K = 64 = 4 x UMMA_K(16)
      const uint64_t a_k = da + static_cast<uint64_t>(k * 2);
      if (p.split_n) {
        umma_bf16<1>(tmem, a_k, make_b(smem_u32(sb)) + static_cast<uint64_t>(k) * p.k_adv16, idesc, k > 0);
        umma_bf16<1>(tmem + 64, a_k, make_b(smem_u32(sb) + p.n_chunk_bytes) + static_cast<uint64_t>(k) * p.k_adv16,
                     idesc, k > 0);
      } else {
        umma_bf16<1>(tmem, a_k, make_b(smem_u32(sb)) + static_cast<uint64_t>(k) * p.k_adv16, idesc, k > 0);
      }
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const int row = warp * 32 + lane;
  for (int c = 0; c < 128; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(t_row + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.D[row * 128 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 128);
  }
}

}  // namespace b200

using namespace b200;

// A [128, 64] bf16 row-major (K contiguous), Bt [64, 128] bf16 row-major (N contiguous) -> D [128, 128] fp32
GLLM_EXPORT int gllm_umma_mn_probe(const void* A, const void* Bt, void* D, int lbo16, int sbo16, int k_adv16,
                                   int n_chunk_bytes, int b_major, int split_n, void* stream) {
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A, 128, 64, 64 * 2, 128, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  if (make_tmap_2d(&tb, Bt, 64, 128, 128 * 2, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  MnProbeParams p;
  p.D = reinterpret_cast<float*>(D);
  p.lbo16 = lbo16; p.sbo16 = sbo16; p.k_adv16 = k_adv16; p.n_chunk_bytes = n_chunk_bytes;
  p.b_major = b_major; p.split_n = split_n;
  constexpr int smem = 32768 + 1024 + 256;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(umma_mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured.done();
  }
  umma_mn_probe_kernel<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(ta, tb, p);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
