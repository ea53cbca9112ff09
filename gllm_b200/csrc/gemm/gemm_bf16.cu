// Persistent, warp-specialised bf16 GEMM for sm_100a:
//
//     C[M, N] = A[M, K] · W[N, K]^T  (+ bias)           (nn.Linear layout, both K-major)
//
//   * operands staged by TMA (128-byte swizzle) through an mbarrier ring,
//   * tcgen05.mma (kind::f16, 128 x BN x 16) issued by one elected thread,
//   * fp32 accumulators double-buffered in TMEM so the epilogue of tile i overlaps the
//     main loop of tile i+1,
//   * four epilogue warps drain TMEM with tcgen05.ld and apply the fused epilogue
//     (bias, SiLU-gate for the merged gate/up projection),
//   * optional fused collectives over NVLink peer memory:
//       - all-gather ⊕ GEMM: the TMA producer gates each M tile on per-row-block
//         "ready" flags that peer ranks set after pushing their activation shard,
//       - GEMM ⊕ reduce-scatter: the epilogue stores each partial tile straight into the
//         owner rank's staging buffer (P2P st.global) and bumps a system-scope counter.
//
// Replaces the reference's cuBLAS F.linear call sites (gllm/layers/linear.py:130,247,339)
// and its GEMM -> NCCL all_reduce sequence (gllm/layers/linear.py:247-250).
#include "../common/host_utils.h"
#include "../common/ptx.cuh"
#include <string.h>

namespace b200 {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle atom row
static constexpr int kUmmaK = 16;
static constexpr int kNumThreads = 192;  // warp0: TMA, warp1: MMA, warps2-5: epilogue
static constexpr int kMaxPeers = 8;

enum Epilogue : int { kEpiStore = 0, kEpiSiluMul = 1 };

struct GemmParams {
  int M, N, K;
  __nv_bfloat16* C;
  int ldc;
  const __nv_bfloat16* bias;
  // all-gather gating (null => disabled)
  const uint32_t* a_ready;      // arrival counters, one per 128-row block of A (written by peers)
  const uint32_t* a_expected;   // device-resident expected counts (CUDA-graph safe), comm/tp_fused.cu
  int m_rot;                    // first M tile to visit (this rank's shard)
  // reduce-scatter push (rs_world == 0 => disabled)
  int rs_world, rs_rank, rows_per_rank;
  uint32_t rs_inc;
  __nv_bfloat16* peer_out[kMaxPeers];
  uint32_t* peer_cnt[kMaxPeers];
  uint32_t prefetch_kb;
  // grouped (MoE) mode: M tile t multiplies the weight slab of expert tile_expert[t]
  const int32_t* tile_expert;
  const int32_t* num_m_tiles_ptr;  // device scalar: number of live M tiles
  int n_per_expert;
  // per-row destination table (EP combine push): row r of C is stored at row_dest[r] (any rank's memory
  // mapped over NVLink); 0 = padding row, not stored. comm/ep_a2a.cu builds the table.
  const int64_t* row_dest;
};

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBudget = 220 * 1024;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // barriers: full[S], empty[S], tmem_full[2], tmem_empty[2] + tmem ptr
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int S = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tmem_full = empty_bar + S;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = p.num_m_tiles_ptr != nullptr ? min(*p.num_m_tiles_ptr, (p.M + kBlockM - 1) / kBlockM)
                                                 : (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_ptr_smem, Cfg::kTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t pf_it = 0;
      int pf_tile = blockIdx.x, pf_kb = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // M tiles are visited starting from this rank's own row shard (m_rot), so an all-gather ⊕ GEMM
        // works on local rows while the peers' rows are still arriving over NVLink
        const int mt = ((tile % num_m) + p.m_rot) % num_m;
        const int m0 = mt * kBlockM;
        const int n0 = (tile / num_m) * BN;
        const int b_row_off = p.tile_expert != nullptr ? p.tile_expert[mt] * p.n_per_expert : 0;
        if (p.a_ready != nullptr) {
          // all-gather ⊕ GEMM: the 128-row block `mt` of A is complete once its arrival counter reached
          // the device-resident expected value (advanced by rs_reduce_norm, comm/tp_fused.cu)
          const uint32_t want = *reinterpret_cast<const volatile uint32_t*>(p.a_expected + mt);
          while (static_cast<int32_t>(ld_acquire_sys(p.a_ready + mt) - want) < 0) {
          }
          asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          // L2 prefetch cursor for the weight operand: runs `p.prefetch_kb` k-blocks ahead of the
          // SMEM ring (across tile boundaries) so weight streaming is not limited by the per-SM
          // number of outstanding DRAM misses of the ring itself.
          while (pf_it < it + p.prefetch_kb && pf_tile < num_tiles) {
            tma_prefetch_l2_2d(&tmap_b, pf_kb * kBlockK, (pf_tile / num_m) * BN);
            ++pf_it;
            if (++pf_kb == num_kb) { pf_kb = 0; pf_tile += gridDim.x; }
          }
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * kBlockK, m0, kEvictNormal);
          tma_load_2d(sb, &tmap_b, &full_bar[s], kb * kBlockK, n0 + b_row_off, kEvictNormal);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
      uint32_t it = 0;
      uint32_t tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(&tmem_empty[buf], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
          const uint64_t da = make_sw128_kmajor_desc(a_addr);
          const uint64_t db = make_sw128_kmajor_desc(b_addr);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 32 B (16 bf16) along K inside the swizzle atom: +2 in 16-byte units
            umma_bf16<1>(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                         (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (((tile % num_m) + p.m_rot) % num_m) * kBlockM;
      const int n0 = (tile / num_m) * BN;
      const uint32_t buf = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(&tmem_full[buf], aph);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_in = row < p.M;
      const uint32_t t_row = tmem_base + buf * BN + (static_cast<uint32_t>(q * 32) << 16);

      // destination row pointer (local C, or the owner rank's staging buffer for RS)
      __nv_bfloat16* crow = nullptr;
      int out_n0 = (EPI == kEpiSiluMul) ? n0 / 2 : n0;
      const int out_N = (EPI == kEpiSiluMul) ? p.N / 2 : p.N;
      bool row_ok = row_in;
      if (row_in && p.row_dest != nullptr) {
        crow = reinterpret_cast<__nv_bfloat16*>(p.row_dest[row]);
        row_ok = crow != nullptr;
      } else if (row_ok) {
        if (p.rs_world > 0) {
          const int owner = row / p.rows_per_rank;
          const int r_local = row - owner * p.rows_per_rank;
          crow = p.peer_out[owner] +
                 (static_cast<size_t>(p.rs_rank) * p.rows_per_rank + r_local) * p.ldc;
        } else {
          crow = p.C + static_cast<size_t>(row) * p.ldc;
        }
      }

      if constexpr (EPI == kEpiStore) {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + c, v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const int col = n0 + c + j;
              if (col < out_N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
                if (p.bias != nullptr) {
                  uint4 bv = *reinterpret_cast<const uint4*>(p.bias + col);
                  float2 b0 = unpack_bf16(bv.x), b1 = unpack_bf16(bv.y), b2 = unpack_bf16(bv.z),
                         b3 = unpack_bf16(bv.w);
                  f[0] += b0.x; f[1] += b0.y; f[2] += b1.x; f[3] += b1.y;
                  f[4] += b2.x; f[5] += b2.y; f[6] += b3.x; f[7] += b3.y;
                }
                uint4 o;
                o.x = pack_bf16(f[0], f[1]);
                o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]);
                o.w = pack_bf16(f[6], f[7]);
                st_v4(crow + col, o);
              }
            }
          }
        }
      } else {
        // SiLU-gate: tile columns [0, BN/2) hold gate, [BN/2, BN) hold up for the same
        // BN/2 output features (weights are interleaved per tile at load time).
        constexpr int H = BN / 2;
#pragma unroll 1
        for (int c = 0; c < H; c += 16) {
          uint32_t g[16], u[16];
          tmem_ld_32x16(t_row + c, g);
          tmem_ld_32x16(t_row + H + c, u);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              const int col = out_n0 + c + j;
              if (col < out_N) {
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float gv = __uint_as_float(g[j + e]);
                  const float uv = __uint_as_float(u[j + e]);
                  f[e] = gv / (1.0f + __expf(-gv)) * uv;
                }
                uint4 o;
                o.x = pack_bf16(f[0], f[1]);
                o.y = pack_bf16(f[2], f[3]);
                o.z = pack_bf16(f[4], f[5]);
                o.w = pack_bf16(f[6], f[7]);
                st_v4(crow + col, o);
              }
            }
          }
        }
      }
      // release the accumulator buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);

      if (p.rs_world > 0) {
        // GEMM ⊕ reduce-scatter: all four epilogue warps have stored their rows; publish the
        // tile to every owner rank whose rows it covers.
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 2 && lane == 0) {
          __threadfence_system();
          const int m1 = min(m0 + kBlockM, p.M);
          const int o0 = m0 / p.rows_per_rank;
          const int o1 = (m1 - 1) / p.rows_per_rank;
          for (int o = o0; o <= o1; ++o) {
            red_add_release_sys(p.peer_cnt[o] + p.rs_rank, p.rs_inc);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool configured = false;
  auto kern = gemm_bf16_kernel<BN, EPI>;
  if (!configured) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::kSmemBytes));
    configured = true;
  }
  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

static int pick_bn(int M, int N, int epi, int forced) {
  if (forced > 0) return forced;
  const int sms = num_sms();
  const int num_m = (M + kBlockM - 1) / kBlockM;
  int best = 128;
  double best_cost = 1e30;
  const int cands[4] = {256, 128, 64, 32};
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (epi == kEpiSiluMul && bn < 64) continue;
    const int tiles = num_m * ((N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    // per-tile cost model: MMA time ∝ bn, plus a fixed per-tile overhead (A traffic, epilogue)
    const double cost = waves * (bn + 48.0);
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace b200

using namespace b200;

// C-ABI entry point. All pointers are device pointers; `stream` is a cudaStream_t.
// a_rows: number of valid rows addressable in A (>= M).  epi: 0 store, 1 SiLU-gate.
// comm: optional pointer to a host-side GemmComm block (may be null).
struct GemmComm {
  const uint32_t* a_ready;
  const uint32_t* a_expected;
  int m_rot;
  int rs_world, rs_rank, rows_per_rank;
  uint32_t rs_inc;
  void* peer_out[kMaxPeers];
  uint32_t* peer_cnt[kMaxPeers];
};

GLLM_EXPORT int gllm_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C,
                               int64_t ldc, int M, int N, int K, const void* bias, int epi,
                               int force_bn, const GemmComm* comm, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % 8) != 0 || (N % 8) != 0 || (lda % 8) != 0 || (ldw % 8) != 0 || (ldc % 8) != 0) {
    fprintf(stderr, "[gllm_b200] gemm_bf16: K, N and leading dims must be multiples of 8\n");
    return 1;
  }
  const int bn = pick_bn(M, N, epi, force_bn);
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, A, M, K, lda * 2, kBlockM, kBlockK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  if (make_tmap_2d(&tb, W, N, K, ldw * 2, bn, kBlockK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = static_cast<int>(ldc);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  {
    static int pf = -1;
    if (pf < 0) {
      const char* e = getenv("GLLM_GEMM_PREFETCH");
      pf = e ? atoi(e) : 0;  // measured: no gain on B200 (profiles/gemm_bf16_v1.md)
    }
    // prefetching only pays when the weights are streamed once (few M tiles share them)
    p.prefetch_kb = (M <= 1024) ? pf : 0;
  }
  if (comm != nullptr) {
    p.a_ready = comm->a_ready;
    p.a_expected = comm->a_expected;
    p.m_rot = comm->m_rot;
    p.rs_world = comm->rs_world;
    p.rs_rank = comm->rs_rank;
    p.rows_per_rank = comm->rows_per_rank;
    p.rs_inc = comm->rs_inc;
    for (int i = 0; i < kMaxPeers; ++i) {
      p.peer_out[i] = reinterpret_cast<__nv_bfloat16*>(comm->peer_out[i]);
      p.peer_cnt[i] = comm->peer_cnt[i];
    }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define DISPATCH(BN_)                                                          \
  case BN_:                                                                    \
    return epi == kEpiSiluMul ? launch_gemm<BN_, kEpiSiluMul>(ta, tb, p, st)   \
                              : launch_gemm<BN_, kEpiStore>(ta, tb, p, st);
  switch (bn) {
    DISPATCH(256)
    DISPATCH(128)
    DISPATCH(64)
    case 32:
      return launch_gemm<32, kEpiStore>(ta, tb, p, st);
    default:
      fprintf(stderr, "[gllm_b200] gemm_bf16: unsupported BN %d\n", bn);
      return 1;
  }
#undef DISPATCH
}

// Number of (m_tile, n_tile) pairs whose rows intersect [row0, row1) — used by the consumer
// of a fused GEMM ⊕ reduce-scatter to know how many tile arrivals to expect per source rank.
GLLM_EXPORT int gllm_gemm_bf16_tiles_covering(int M, int N, int epi, int force_bn, int row0,
                                              int row1) {
  if (row1 > M) row1 = M;
  if (row1 <= row0) return 0;
  const int bn = pick_bn(M, N, epi, force_bn);
  const int t0 = row0 / kBlockM;
  const int t1 = (row1 - 1) / kBlockM;
  return (t1 - t0 + 1) * ((N + bn - 1) / bn);
}

// Grouped (MoE) GEMM: rows of A are expert-sorted and padded to 128-row tiles; tile t uses the weight
// slab W[tile_expert[t]] ([E, N, K] contiguous). The live tile count is read on the device.
// epi: 0 store, 1 SiLU-gate (slab rows interleaved per 64 like the dense gate/up weight).
GLLM_EXPORT int gllm_moe_grouped_gemm(const void* A, int64_t lda, const void* W, void* C, int64_t ldc,
                                      int max_tiles, int N, int K, int E, const void* tile_expert,
                                      const void* num_tiles_ptr, int epi, const void* row_dest, void* stream) {
  if (max_tiles <= 0) return 0;
  if ((K % 8) != 0 || (N % 8) != 0 || (lda % 8) != 0 || (ldc % 8) != 0) return 1;
  const int bn = 128;  // 64|64 gate/up interleave for the SiLU epilogue; good balance for expert tiles
  CUtensorMap ta, tb;
  const int M = max_tiles * kBlockM;
  if (make_tmap_2d(&ta, A, M, K, lda * 2, kBlockM, kBlockK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  if (make_tmap_2d(&tb, W, static_cast<uint64_t>(E) * N, K, static_cast<uint64_t>(K) * 2, bn, kBlockK,
                   CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.C = reinterpret_cast<__nv_bfloat16*>(C);
  p.ldc = static_cast<int>(ldc);
  p.tile_expert = reinterpret_cast<const int32_t*>(tile_expert);
  p.num_m_tiles_ptr = reinterpret_cast<const int32_t*>(num_tiles_ptr);
  p.n_per_expert = N;
  p.row_dest = reinterpret_cast<const int64_t*>(row_dest);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return epi == kEpiSiluMul ? launch_gemm<128, kEpiSiluMul>(ta, tb, p, st) : launch_gemm<128, kEpiStore>(ta, tb, p, st);
}
