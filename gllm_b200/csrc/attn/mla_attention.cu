// Multi-head latent attention (DeepSeek V2 / V3 / R1, Kimi) over the paged LATENT cache, absorbed form:
//
//   q_full[t, h, :] = [ q_nope[t, h] · W_UK[h]  (512) | rope(q_pe[t, h]) (64) ]          (built by the caller)
//   s[t, h, j]      = q_full[t, h] · latent[j]            latent[j] = [ c_kv[j] (512) | rope(k_pe[j]) (64) ]
//   out_lat[t, h]   = softmax_j(s * scale) · c_kv[j]      (512)          -> caller applies W_UV[h]
//
// i.e. multi-query attention with one shared 576-wide key whose first 512 columns are also the value, so a KV
// tile is fetched ONCE (TMA page boxes, 128-byte swizzle, mbarrier ring, producer warp) and serves both GEMMs.
// This replaces the reference's Triton split-KV MQA decode kernels and its gather/up-project/FlashAttention
// prefill loop (gllm/layers/attention.py:129-391, gllm/layers/ops/triton_decode_attention.py).
//
// One CTA = 16 query heads of one token x one KV split. The four math warps split each 64-key tile for
// S = Q K^T (16 keys per warp, k = 576 from a padded smem copy of Q), exchange the row maxima and the bf16
// probabilities through shared memory, then split the VALUE columns for O += P V (128 of the 512 columns per
// warp, 64 fp32 accumulators per thread). Every token is its own "sequence" with kv_len = position + 1, so the
// same kernel serves decode and (chunked / prefix-cached) prefill; splits are merged by attn_merge_kernel.
//
// Latent cache layout: [num_pages, 1, 9, page_size, 64] bf16 (the generic paged layout with Hkv = 1, D = 576).
#include "attn_common.cuh"

namespace b200 {

static constexpr int kLat = 576;        // kv_lora_rank (512) + qk_rope_head_dim (64)
static constexpr int kLatV = 512;       // value width = kv_lora_rank
static constexpr int kMlaStages = 2;
static constexpr int kMlaTileBytes = kTileN * kLat * 2;                 // 73,728
static constexpr int kQStride = kLat * 2 + 16;                          // padded row pitch of Q in smem (bytes)
static constexpr int kPStride = kTileN * 2 + 16;                        // padded row pitch of P in smem (bytes)
static constexpr int kMlaSmem = kMlaStages * kMlaTileBytes + 16 * kQStride + 16 * kPStride + 4 * 16 * 4 * 2 + 1024 + 128;

struct MlaParams {
  const __nv_bfloat16* q;     // [tokens, H, 576]
  __nv_bfloat16* out;         // [tokens, H, 512]
  float* part_o;              // [tokens, H, splits, 512]
  float* part_lse;            // [tokens, H, splits]
  const int32_t* block_table; // [seqs, max_blocks]
  const int32_t* tok_seq;     // [tokens] sequence (block-table row) of every token; null => token index
  const int32_t* positions;   // [tokens] absolute position of the token (kv_len = position + 1)
  int max_blocks, H, page_size, num_splits;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads)
mla_attn_kernel(const __grid_constant__ CUtensorMap tmap, const MlaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem + kMlaStages * kMlaTileBytes;
  uint8_t* p_smem = q_smem + 16 * kQStride;
  float* red_max = reinterpret_cast<float*>(p_smem + 16 * kPStride);   // [4 warps][16 rows]
  float* red_sum = red_max + 4 * 16;                                    // [4 warps][16 rows]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red_sum + 4 * 16);
  uint64_t* empty_bar = full_bar + kMlaStages;

  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hbase = blockIdx.x * 16;
  const int tok = blockIdx.y;
  const int split = blockIdx.z;
  const int seq = p.tok_seq != nullptr ? p.tok_seq[tok] : tok;
  const int kv_len = p.positions[tok] + 1;
  const int n_tiles = (kv_len + kTileN - 1) / kTileN;
  const int tps = (n_tiles + p.num_splits - 1) / p.num_splits;
  const int tile_begin = min(split * tps, n_tiles);
  const int tile_end = min(tile_begin + tps, n_tiles);
  const int page_bytes = p.page_size * kLat * 2;
  const int n_rows = min(16, p.H - hbase);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMlaStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kMathWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kMathWarps) {
    // ---------------- producer: one TMA box per page, K == V ----------------
    if (lane == 0 && tile_end > tile_begin) {
      const int32_t* bt = p.block_table + static_cast<size_t>(seq) * p.max_blocks;
      const int pages_per_tile = kTileN / p.page_size;
      const int last_page = (kv_len - 1) / p.page_size;
      uint32_t it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it % kMlaStages;
        mbar_wait(&empty_bar[s], ((it / kMlaStages) & 1) ^ 1);
        uint8_t* sk = smem + s * kMlaTileBytes;
        mbar_expect_tx(&full_bar[s], kMlaTileBytes);
        for (int j = 0; j < pages_per_tile; ++j) {
          int pi = tile * pages_per_tile + j;
          if (pi > last_page) pi = last_page;  // keep smem finite for masked columns
          tma_load_3d(sk + j * page_bytes, &tmap, &full_bar[s], 0, 0, bt[pi] * (kLat / 64));
        }
      }
    }
    return;
  }

  // ---------------- math warps ----------------
  const int g = lane >> 2, t = lane & 3;
  // Q -> smem (rows beyond the head count are zero)
  {
    const __nv_bfloat16* qb = p.q + (static_cast<size_t>(tok) * p.H + hbase) * kLat;
    for (int i = threadIdx.x; i < 16 * (kLat / 8); i += kMathWarps * 32) {
      const int r = i / (kLat / 8), c = i % (kLat / 8);
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < n_rows) v = *reinterpret_cast<const uint4*>(qb + static_cast<size_t>(r) * kLat + c * 8);
      *reinterpret_cast<uint4*>(q_smem + r * kQStride + c * 16) = v;
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const uint32_t q_base = smem_u32(q_smem);
  const uint32_t p_base = smem_u32(p_smem);

  float o[16][4];  // this warp's 128 value columns [128 * warp, +128): 16 n-tiles of 8
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};  // partial row sums over THIS warp's keys (all warps share the same m_run)

  uint32_t it = 0;
  for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
    const int s = it % kMlaStages;
    mbar_wait(&full_bar[s], (it / kMlaStages) & 1);
    const uint32_t sk = smem_u32(smem + s * kMlaTileBytes);
    const int tok0 = warp * 16;  // this warp's 16 keys inside the tile

    // S = Q K^T for 16 rows x 16 keys, k = 576
    float sc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll 4
    for (int ks = 0; ks < kLat / 16; ks += 2) {
      uint32_t qa[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h)
        ldsm_x4(q_base + (lane & 15) * kQStride + ((ks + h) * 16 + (lane >> 4) * 8) * 2, qa[h][0], qa[h][1], qa[h][2],
                qa[h][3]);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        uint32_t b0, b1, b2, b3;
        const int key = tok0 + nt * 8 + (lane & 7);
        const int c8 = ks * 2 + (lane >> 3);
        ldsm_x4(sk + tile_off(key, c8, p.page_size, page_bytes), b0, b1, b2, b3);
        mma_bf16_16816(sc[nt], qa[0], b0, b1);
        mma_bf16_16816(sc[nt], qa[1], b2, b3);
      }
    }
    // mask, local row max over this warp's 16 keys
    const int abs0 = tile * kTileN + tok0 + 2 * t;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = abs0 + nt * 8 + (e & 1);
        float v = sc[nt][e] * p.scale_log2;
        if (col >= kv_len) v = -INFINITY;
        sc[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    if (t == 0) {
      red_max[warp * 16 + g] = mx[0];
      red_max[warp * 16 + g + 8] = mx[1];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    float alpha[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = g + r * 8;
      float m_tile = fmaxf(fmaxf(red_max[row], red_max[16 + row]), fmaxf(red_max[32 + row], red_max[48 + row]));
      const float m_new = fmaxf(m_run[r], m_tile);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[r] = exp2f(m_run[r] - m_use[r]);
      m_run[r] = m_new;
      l_run[r] *= alpha[r];
    }
    // P (bf16) for this warp's 16 keys -> smem [16 rows][64 keys]
    {
      float pv[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pv[nt][e] = exp2f(sc[nt][e] - m_use[e >> 1]);
          l_run[e >> 1] += pv[nt][e];
        }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = tok0 + nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(p_smem + g * kPStride + col * 2) = pack_bf16(pv[nt][0], pv[nt][1]);
        *reinterpret_cast<uint32_t*>(p_smem + (g + 8) * kPStride + col * 2) = pack_bf16(pv[nt][2], pv[nt][3]);
      }
    }
#pragma unroll
    for (int nd = 0; nd < 16; ++nd) {
      o[nd][0] *= alpha[0]; o[nd][1] *= alpha[0];
      o[nd][2] *= alpha[1]; o[nd][3] *= alpha[1];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // O[:, 128 * warp + ...] += P (16 x 64) * V (64 keys x 128 columns of the same tile)
#pragma unroll
    for (int kk = 0; kk < kTileN / 16; ++kk) {
      uint32_t pa[4];
      ldsm_x4(p_base + (lane & 15) * kPStride + (kk * 16 + (lane >> 4) * 8) * 2, pa[0], pa[1], pa[2], pa[3]);
#pragma unroll
      for (int nd = 0; nd < 16; nd += 2) {
        uint32_t b0, b1, b2, b3;
        const int key = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int c8 = warp * 16 + nd + (lane >> 4);
        ldsm_x4_t(sk + tile_off(key, c8, p.page_size, page_bytes), b0, b1, b2, b3);
        mma_bf16_16816(o[nd], pa, b0, b1);
        mma_bf16_16816(o[nd + 1], pa, b2, b3);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // ---------------- finish: total row sums across the warps, write this warp's value columns ----------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // red_max no longer read by anybody
  if (t == 0) {
    red_sum[warp * 16 + g] = l_run[0];
    red_sum[warp * 16 + g + 8] = l_run[1];
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = g + r * 8;
    if (row >= n_rows) continue;
    const float l_all = red_sum[row] + red_sum[16 + row] + red_sum[32 + row] + red_sum[48 + row];
    const float inv = l_all > 0.f ? 1.f / l_all : 0.f;
    const size_t th = static_cast<size_t>(tok) * p.H + hbase + row;
#pragma unroll
    for (int nd = 0; nd < 16; ++nd) {
      const int c = warp * 128 + nd * 8 + 2 * t;
      const float a = o[nd][2 * r] * inv, b = o[nd][2 * r + 1] * inv;
      if (p.num_splits == 1) {
        *reinterpret_cast<uint32_t*>(p.out + th * kLatV + c) = pack_bf16(a, b);
      } else {
        *reinterpret_cast<float2*>(p.part_o + (th * p.num_splits + split) * kLatV + c) = make_float2(a, b);
      }
    }
    if (p.num_splits > 1 && warp == 0 && t == 0) {
      p.part_lse[th * p.num_splits + split] = (l_all > 0.f) ? (m_run[r] + log2f(l_all)) : -INFINITY;
    }
  }
}

// Per token: interleaved-pair RoPE on the rope dims of every query head (written into q_full[:, :, 512:576]) and
// of the shared key, then the latent row [c_kv (512) | rope(k_pe) (64)] goes into its paged-cache slot.
// (reference: rotary_embedding + concat_and_cache_mla, gllm/layers/attention.py:443-451)
__global__ void mla_rope_cache_kernel(const __nv_bfloat16* __restrict__ q_pe, int64_t q_ts, int64_t q_hs, int H,
                                      __nv_bfloat16* __restrict__ q_full, const __nv_bfloat16* __restrict__ k_pe,
                                      int64_t k_ts, const __nv_bfloat16* __restrict__ kv_c, int64_t c_ts,
                                      const float* __restrict__ cos_sin, const int32_t* __restrict__ positions,
                                      const int32_t* __restrict__ slots, __nv_bfloat16* __restrict__ cache,
                                      int page_size) {
  griddep_launch();
  griddep_wait();
  const int tok = blockIdx.x;
  const int pos = positions[tok];
  const float* cs = cos_sin + static_cast<size_t>(pos) * 64;  // [cos (32) | sin (32)]
  for (int i = threadIdx.x; i < H * 32; i += blockDim.x) {
    const int h = i >> 5, pr = i & 31;
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(q_pe + tok * q_ts + h * q_hs + 2 * pr);
    const float2 x = unpack_bf16(raw);
    const float c = cs[pr], sn = cs[32 + pr];
    *reinterpret_cast<uint32_t*>(q_full + (static_cast<size_t>(tok) * H + h) * kLat + kLatV + 2 * pr) =
        pack_bf16(x.x * c - x.y * sn, x.y * c + x.x * sn);
  }
  const int slot = slots != nullptr ? slots[tok] : -1;
  if (slot < 0) return;
  const int page = slot / page_size, off = slot - page * page_size;
  __nv_bfloat16* dst = cache + (static_cast<size_t>(page) * (kLat / 64) * page_size + off) * 64;
  const size_t slab_stride = static_cast<size_t>(page_size) * 64;
  for (int i = threadIdx.x; i < kLatV / 8; i += blockDim.x) {  // 64 chunks of 8 bf16
    const uint4 v = *reinterpret_cast<const uint4*>(kv_c + tok * c_ts + i * 8);
    *reinterpret_cast<uint4*>(dst + (i >> 3) * slab_stride + (i & 7) * 8) = v;
  }
  if (threadIdx.x < 32) {
    const int pr = threadIdx.x;
    const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(k_pe + tok * k_ts + 2 * pr));
    const float c = cs[pr], sn = cs[32 + pr];
    *reinterpret_cast<uint32_t*>(dst + 8 * slab_stride + 2 * pr) = pack_bf16(x.x * c - x.y * sn, x.y * c + x.x * sn);
  }
}

// defined in paged_attention.cu
int launch_attn_merge(const float* part_o, const float* part_lse, __nv_bfloat16* out, int rows, int num_splits, int D,
                      cudaStream_t st);

}  // namespace b200

using namespace b200;

// q [tokens, H, 576] bf16 contiguous -> out [tokens, H, 512]; cache = latent pages [num_pages, 1, 9, page, 64].
// part_o / part_lse are only touched when num_splits > 1.
GLLM_EXPORT int gllm_mla_attention(const void* q, void* out, const void* cache, int64_t num_pages,
                                   const void* block_table, const void* tok_seq, const void* positions, void* part_o,
                                   void* part_lse, int tokens, int max_blocks, int H, int page_size, int num_splits,
                                   float scale, void* stream) {
  if (tokens <= 0) return 0;
  if (page_size < 8 || kTileN % page_size != 0) {
    fprintf(stderr, "[gllm_b200] mla_attention: unsupported page_size=%d\n", page_size);
    return 1;
  }
  CUtensorMap tm;
  if (get_kv_tmap(cache, num_pages, 1, kLat, page_size, &tm)) return 1;
  MlaParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.part_o = reinterpret_cast<float*>(part_o);
  p.part_lse = reinterpret_cast<float*>(part_lse);
  p.block_table = reinterpret_cast<const int32_t*>(block_table);
  p.tok_seq = reinterpret_cast<const int32_t*>(tok_seq);
  p.positions = reinterpret_cast<const int32_t*>(positions);
  p.max_blocks = max_blocks; p.H = H; p.page_size = page_size;
  p.num_splits = num_splits < 1 ? 1 : num_splits;
  p.scale_log2 = scale * 1.4426950408889634f;
  static PerDeviceOnce configured;
  if (configured.need()) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(mla_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMlaSmem));
    configured.done();
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((H + 15) / 16, tokens, p.num_splits);
  CUDA_CHECK_RET(launch_pdl(mla_attn_kernel, grid, dim3(kAttnThreads), kMlaSmem, st, tm, p));
  if (p.num_splits > 1) {
    if (launch_attn_merge(p.part_o, p.part_lse, p.out, tokens * H, p.num_splits, kLatV, st)) return 1;
  }
  return 0;
}

// q_pe: strided view [T, H, 64] of the projected queries; k_pe [T, 64]; kv_c [T, 512] (normalised latent);
// cos_sin fp32 [max_pos, 64]; writes q_full[:, :, 512:] and the cache rows at `slots`.
GLLM_EXPORT int gllm_mla_rope_cache(const void* q_pe, int64_t q_ts, int64_t q_hs, int H, void* q_full,
                                    const void* k_pe, int64_t k_ts, const void* kv_c, int64_t c_ts,
                                    const void* cos_sin, const void* positions, const void* slots, void* cache,
                                    int page_size, int tokens, void* stream) {
  if (tokens <= 0) return 0;
  CUDA_CHECK_RET(launch_pdl(mla_rope_cache_kernel, dim3(tokens), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream),
                            reinterpret_cast<const __nv_bfloat16*>(q_pe), q_ts, q_hs, H,
                            reinterpret_cast<__nv_bfloat16*>(q_full), reinterpret_cast<const __nv_bfloat16*>(k_pe), k_ts,
                            reinterpret_cast<const __nv_bfloat16*>(kv_c), c_ts, reinterpret_cast<const float*>(cos_sin),
                            reinterpret_cast<const int32_t*>(positions), reinterpret_cast<const int32_t*>(slots),
                            reinterpret_cast<__nv_bfloat16*>(cache), page_size));
  return 0;
}
