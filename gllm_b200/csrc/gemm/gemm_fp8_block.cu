// Block-scaled FP8 (e4m3) GEMM for sm_100a, DeepSeek-V3 / Qwen3-FP8 checkpoint semantics
// (reference: gllm/layers/quantization/fp8.py:54-250 — Triton w8a8_block_fp8_matmul):
//
//   C[M,N] = sum_kb  (A8[M, kb] · W8[N, kb]^T) * a_s[M, kb] * w_s[N/128, kb]      (+ bias) -> bf16
//   A8: activations quantised per token per 128-wide K group (dynamic, amax/448), a_s fp32 [K/128, M]
//       (stored K-block-major so a warp's 32 rows read 32 consecutive floats)
//   W8: weights e4m3 [N, K] with fp32 scales per 128x128 block, w_s [N/128, K/128]
//
// The checkpoint scales are arbitrary fp32 values (not the power-of-two UE8M0 factors that
// `kind::mxf8f6f4.block_scale` applies in hardware), so each 128-deep K block is multiplied by
// `tcgen05.mma.kind::f8f6f4` into its own TMEM buffer (4 MMAs of K=32, the first one overwriting)
// and the epilogue warps promote it: acc += partial * a_s[row] * w_s[tile] in fp32 registers. Two
// TMEM buffers ping-pong per K block so the tensor core runs block kb+1 while the CUDA cores fold
// block kb. Same TMA (SWIZZLE_128B, 128 fp8 = one 128-byte row) / mbarrier ring as gemm_bf16.cu.
//
// Also here: the dynamic per-token-group activation quantiser (reference fp8.py:354-552).
#include <cuda_fp8.h>
#include <string.h>

#include "../common/host_utils.h"
#include "../common/ptx.cuh"

namespace b200 {

static constexpr int kFBM = 128, kFBN = 128, kFBK = 128;  // K block = 128 fp8 = 128 B
static constexpr int kFStages = 6;
static constexpr int kFThreads = 192;

struct Fp8Params {
  int M, N, K;
  __nv_bfloat16* C;
  int ldc;
  const float* a_s;  // [K/128, M]
  int lda_s;         // = M (row pitch of a_s)
  const float* w_s;  // [N/128, K/128]  (grouped: [E][N/64][K/128], one scale row per 64 weight rows)
  const __nv_bfloat16* bias;
  // grouped (MoE) mode: M tile t multiplies the e4m3 slab of expert tile_expert[t]; the weight scales are
  // stored per 64-row half tile so a [64 gate | 64 up] interleaved tile can carry two different block scales
  const int32_t* tile_expert;
  const int32_t* num_m_tiles_ptr;
  int n_per_expert;      // rows of one expert slab
  int64_t ws_stride_e;   // floats per expert in w_s
  int silu;              // 1: out[:, j] = silu(acc[j]) * acc[64 + j]  (64 output columns per tile)
};

__global__ void __launch_bounds__(kFThreads, 1)
gemm_fp8_block_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const Fp8Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kABytes = kFBM * kFBK, kBBytes = kFBN * kFBK, kStageBytes = kABytes + kBBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kFStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kFStages;
  uint64_t* tmem_full = empty_bar + kFStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool grouped = p.tile_expert != nullptr;
  const int num_m = p.num_m_tiles_ptr != nullptr ? min(*p.num_m_tiles_ptr, (p.M + kFBM - 1) / kFBM)
                                                 : (p.M + kFBM - 1) / kFBM;
  const int num_n = (p.N + kFBN - 1) / kFBN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / kFBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < kFStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % num_m) * kFBM, n0 = (tile / num_m) * kFBN;
        const int b_row_off = grouped ? p.tile_expert[tile % num_m] * p.n_per_expert : 0;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kFStages;
          mbar_wait(&empty_bar[s], ((it / kFStages) & 1) ^ 1);
          uint8_t* sa = smem + s * kStageBytes;
          mbar_expect_tx(&full_bar[s], kStageBytes);
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kFBK, m0, kEvictNormal);
          tma_load_2d(sa + kABytes, &tmap_b, &full_bar[s], kb * kFBK, n0 + b_row_off, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_e4m3(kFBM, kFBN);
      uint32_t it = 0;  // global k-block counter == TMEM hand-off counter
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kFStages;
          const uint32_t buf = it & 1;
          mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1);
          mbar_wait(&full_bar[s], (it / kFStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
          const uint64_t da = make_sw128_kmajor_desc(a_addr);
          const uint64_t db = make_sw128_kmajor_desc(a_addr + kABytes);
          const uint32_t d_tmem = tmem_base + buf * kFBN;
#pragma unroll
          for (int k = 0; k < kFBK / 32; ++k)  // UMMA K = 32 for 8-bit operands: +32 B per step
            umma_f8(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k > 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          umma_commit(&tmem_full[buf]);
        }
      }
    }
  } else {
    const int q = warp & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * kFBM, n0 = (tile / num_m) * kFBN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      float acc[kFBN];
#pragma unroll
      for (int i = 0; i < kFBN; ++i) acc[i] = 0.f;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t buf = it & 1;
        const float sa = row_ok ? p.a_s[static_cast<size_t>(kb) * p.lda_s + row] : 0.f;
        float sc, sc_hi;
        if (grouped) {
          const float* ws = p.w_s + static_cast<size_t>(p.tile_expert[tile % num_m]) * p.ws_stride_e +
                            static_cast<size_t>(n0 / 64) * num_kb + kb;
          sc = sa * ws[0];
          sc_hi = sa * ws[num_kb];
        } else {
          sc = sc_hi = sa * p.w_s[static_cast<size_t>(n0 / kFBN) * num_kb + kb];
        }
        mbar_wait(&tmem_full[buf], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + buf * kFBN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < kFBN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[c + j] = fmaf(__uint_as_float(v[j]), c < 64 ? sc : sc_hi, acc[c + j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      if (row_ok && p.silu) {
        __nv_bfloat16* crow = p.C + static_cast<size_t>(row) * p.ldc;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          const int col = n0 / 2 + c;
          if (col < p.N / 2) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float gv = acc[c + e];
              f[e] = gv / (1.0f + __expf(-gv)) * acc[64 + c + e];
            }
            st_v4(crow + col, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                         pack_bf16(f[6], f[7])));
          }
        }
      } else if (row_ok) {
        __nv_bfloat16* crow = p.C + static_cast<size_t>(row) * p.ldc;
#pragma unroll
        for (int c = 0; c < kFBN; c += 8) {
          const int col = n0 + c;
          if (col < p.N) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = acc[c + e];
            if (p.bias != nullptr) {
              const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + col);
              const float2 b0 = unpack_bf16(bv.x), b1 = unpack_bf16(bv.y), b2 = unpack_bf16(bv.z), b3 = unpack_bf16(bv.w);
              f[0] += b0.x; f[1] += b0.y; f[2] += b1.x; f[3] += b1.y; f[4] += b2.x; f[5] += b2.y; f[6] += b3.x; f[7] += b3.y;
            }
            st_v4(crow + col, make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                         pack_bf16(f[6], f[7])));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 256);
  }
}

// dynamic per-token-group (128) quantisation: x bf16 [M,K] -> q e4m3 [M,K], scales fp32 [K/128, M]
__global__ void fp8_quant_group_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_fp8_e4m3* __restrict__ q,
                                       float* __restrict__ scales, int M, int K) {
  // one warp per (row, group): 32 lanes x 4 elements = 128
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = K / 128;
  if (wid >= M * groups) return;
  const int row = wid / groups, g = wid % groups;
  const __nv_bfloat16* src = x + static_cast<size_t>(row) * ldx + g * 128 + lane * 4;
  const uint2 raw = *reinterpret_cast<const uint2*>(src);
  const float2 a = unpack_bf16(raw.x), b = unpack_bf16(raw.y);
  float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(b.x), fabsf(b.y)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  amax = fmaxf(amax, 1e-10f);
  const float scale = amax / 448.f;
  const float inv = 1.f / scale;
  __nv_fp8_e4m3 o4[4] = {__nv_fp8_e4m3(a.x * inv), __nv_fp8_e4m3(a.y * inv), __nv_fp8_e4m3(b.x * inv),
                         __nv_fp8_e4m3(b.y * inv)};
  *reinterpret_cast<uint32_t*>(q + static_cast<size_t>(row) * K + g * 128 + lane * 4) = *reinterpret_cast<uint32_t*>(o4);
  if (lane == 0) scales[static_cast<size_t>(g) * M + row] = scale;
}

}  // namespace b200

using namespace b200;

GLLM_EXPORT int gllm_fp8_quant_group(const void* x, int64_t ldx, void* q, void* scales, int M, int K, void* stream) {
  if (M <= 0) return 0;
  if (K % 128 != 0) return 1;
  const int warps = M * (K / 128);
  fp8_quant_group_kernel<<<(warps * 32 + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_fp8_e4m3*>(q),
      reinterpret_cast<float*>(scales), M, K);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// A8 [M,K] e4m3 contiguous, a_s [K/128, M]; W8 [N,K] e4m3 contiguous, w_s [ceil(N/128), K/128]
GLLM_EXPORT int gllm_gemm_fp8_block(const void* A8, const void* a_s, const void* W8, const void* w_s, void* C,
                                    int64_t ldc, int M, int N, int K, const void* bias, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K % 128 != 0 || N % 8 != 0) {
    fprintf(stderr, "[gllm_b200] gemm_fp8_block: K %% 128 and N %% 8 required\n");
    return 1;
  }
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A8, M, K, K, kFBM, kFBK, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return 1;
  if (make_tmap_2d(&tb, W8, N, K, K, kFBN, kFBK, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return 1;
  Fp8Params p;
  p.M = M; p.N = N; p.K = K;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = static_cast<int>(ldc);
  p.a_s = reinterpret_cast<const float*>(a_s);
  p.lda_s = M;
  p.w_s = reinterpret_cast<const float*>(w_s);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  p.tile_expert = nullptr; p.num_m_tiles_ptr = nullptr; p.n_per_expert = 0; p.ws_stride_e = 0; p.silu = 0;
  constexpr int smem_bytes = kFStages * (kFBM * kFBK + kFBN * kFBK) + 1024 + 256;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_fp8_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured.done();
  }
  const int tiles = ((M + kFBM - 1) / kFBM) * ((N + kFBN - 1) / kFBN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_fp8_block_kernel<<<grid, kFThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(ta, tb, p);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// Grouped (MoE) block-scaled fp8 GEMM: A8 [max_tiles*128, K] e4m3 (expert-sorted, 128-row tiles), a_s
// [K/128, max_tiles*128]; W8 [E, N, K] e4m3, w_s [E, N/64, K/128] (one scale row per 64 weight rows);
// epi 1 = SiLU gate on [64 gate | 64 up] interleaved tiles -> C [rows, N/2].
GLLM_EXPORT int gllm_moe_grouped_gemm_fp8(const void* A8, const void* a_s, const void* W8, const void* w_s, void* C,
                                          int64_t ldc, int max_tiles, int N, int K, int E, const void* tile_expert,
                                          const void* num_tiles_ptr, int epi, void* stream) {
  if (max_tiles <= 0) return 0;
  if (K % 128 != 0 || N % 128 != 0) {
    fprintf(stderr, "[gllm_b200] moe_grouped_gemm_fp8: K %% 128 and N %% 128 required\n");
    return 1;
  }
  const int M = max_tiles * kFBM;
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A8, M, K, K, kFBM, kFBK, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return 1;
  if (make_tmap_2d(&tb, W8, static_cast<uint64_t>(E) * N, K, K, kFBN, kFBK, CU_TENSOR_MAP_DATA_TYPE_UINT8)) return 1;
  Fp8Params p;
  p.M = M; p.N = N; p.K = K;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = static_cast<int>(ldc);
  p.a_s = reinterpret_cast<const float*>(a_s);
  p.lda_s = M;
  p.w_s = reinterpret_cast<const float*>(w_s);
  p.bias = nullptr;
  p.tile_expert = reinterpret_cast<const int32_t*>(tile_expert);
  p.num_m_tiles_ptr = reinterpret_cast<const int32_t*>(num_tiles_ptr);
  p.n_per_expert = N;
  p.ws_stride_e = static_cast<int64_t>(N / 64) * (K / 128);
  p.silu = epi == 1 ? 1 : 0;
  constexpr int smem_bytes = kFStages * (kFBM * kFBK + kFBN * kFBK) + 1024 + 256;
  CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_fp8_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int tiles = max_tiles * (N / kFBN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_fp8_block_kernel<<<grid, kFThreads, smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(ta, tb, p);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
