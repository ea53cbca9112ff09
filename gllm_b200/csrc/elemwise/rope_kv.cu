// Fused  [per-head q/k RMSNorm] -> RoPE (NeoX / GPT-J / partial / M-RoPE) -> paged KV-cache write.
//
// One pass over the QKV projection output replaces four kernels of the reference:
// q_norm/k_norm (gllm/models/qwen3.py:75-80), `_C.rotary_embedding`
// (gllm/layers/rotary_embedding.py:91-110), the Triton M-RoPE kernel (:342-506) and
// `reshape_and_cache_flash` (gllm/memory_manager.py:154-163).
//
// KV cache layout (ours, chosen for TMA): [num_pages, Hkv, D/64, page_size, 64] bf16 — each
// (page, head, 64-wide slice) is a contiguous page_size x 128 B slab that one TMA box with the
// 128-byte swizzle turns into a conflict-free tile for the attention kernels.
#include "../common/host_utils.h"
#include "../common/ptx.cuh"

namespace b200 {

struct RopeParams {
  __nv_bfloat16* q; int64_t q_ts, q_hs; int Hq;
  __nv_bfloat16* k; int64_t k_ts, k_hs; int Hkv;
  const __nv_bfloat16* v; int64_t v_ts, v_hs;
  const __nv_bfloat16* q_norm_w;
  const __nv_bfloat16* k_norm_w;
  const float* cos_sin;  // [max_pos, rot]: cos | sin
  int D, rot, neox, T;
  const int32_t* positions;  // [T] or [3, T] (mrope)
  const int32_t* slots;      // [T] or null
  float eps;
  __nv_bfloat16* k_cache;
  __nv_bfloat16* v_cache;
  int sec0, sec1, page_size;  // sec0 > 0: chunked M-RoPE [T|H|W]; sec0 < 0: interleaved (H = -sec0, W = sec1)
  int64_t pos_stride;         // row stride of a [3, T] positions tensor
};

// which of the (t, h, w) position rows rotary pair i uses
__device__ __forceinline__ int mrope_axis(int i, int sec0, int sec1) {
  if (sec0 > 0) return i < sec0 ? 0 : (i < sec0 + sec1 ? 1 : 2);
  if (sec0 < 0) {
    const int r = i % 3;
    if (r == 1 && i < -sec0 * 3) return 1;
    if (r == 2 && i < sec1 * 3) return 2;
  }
  return 0;
}

template <int C>  // elements per lane, D = 32 * C
__global__ void rope_kv_kernel(const RopeParams p) {
  griddep_launch();
  griddep_wait();
  const int t = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int D = 32 * C;
  const int half = p.rot >> 1;
  const int e0 = lane * C;  // first element of this lane
  const bool do_rope = p.rot > 0;
  const int n_heads = p.Hq + p.Hkv + ((p.v != nullptr && p.slots != nullptr) ? p.Hkv : 0);

  int pos[3];
  if (p.sec0 != 0) {
    pos[0] = p.positions[t]; pos[1] = p.positions[p.pos_stride + t]; pos[2] = p.positions[2 * p.pos_stride + t];
  } else {
    pos[0] = pos[1] = pos[2] = p.positions != nullptr ? p.positions[t] : 0;
  }
  const int slot = p.slots != nullptr ? p.slots[t] : -1;

  // one head per warp: blockIdx.y strides the head list so a decode batch still fills the machine and no
  // warp walks several heads serially (the per-head chain of dependent global loads is the whole cost)
  for (int hh = blockIdx.y * nwarps + warp; hh < n_heads; hh += nwarps * gridDim.y) {
    const bool is_q = hh < p.Hq;
    const bool is_k = !is_q && hh < p.Hq + p.Hkv;
    const int h = is_q ? hh : (is_k ? hh - p.Hq : hh - p.Hq - p.Hkv);
    float x[C];
    if (is_q || is_k) {
      __nv_bfloat16* ptr = is_q ? p.q + t * p.q_ts + h * p.q_hs : p.k + t * p.k_ts + h * p.k_hs;
      const __nv_bfloat16* nw = is_q ? p.q_norm_w : p.k_norm_w;
#pragma unroll
      for (int e = 0; e < C; e += 2) {
        float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(ptr + e0 + e));
        x[e] = f.x; x[e + 1] = f.y;
      }
      if (nw != nullptr) {
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < C; ++e) ss += x[e] * x[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = rsqrtf(ss / static_cast<float>(D) + p.eps);
#pragma unroll
        for (int e = 0; e < C; e += 2) {
          float2 wv = unpack_bf16(*reinterpret_cast<const uint32_t*>(nw + e0 + e));
          // round to bf16 after the norm like the unfused reference does
          x[e] = __bfloat162float(__float2bfloat16(x[e] * inv * wv.x));
          x[e + 1] = __bfloat162float(__float2bfloat16(x[e + 1] * inv * wv.y));
        }
      }
      if (do_rope) {
        if (p.neox) {
          // partner element lives `half` elements away -> lane +- half / C
          const int dl = half / C;
          const bool lo = e0 < half;
          const bool in_rot = e0 < p.rot;
          const int src = lo ? lane + dl : lane - dl;
          float y[C];
#pragma unroll
          for (int e = 0; e < C; ++e) y[e] = __shfl_sync(0xffffffffu, x[e], src & 31);
          if (in_rot) {
#pragma unroll
            for (int e = 0; e < C; ++e) {
              const int i = (e0 + e) % half;
              const int ps = pos[mrope_axis(i, p.sec0, p.sec1)];
              const float c = p.cos_sin[static_cast<size_t>(ps) * p.rot + i];
              const float s = p.cos_sin[static_cast<size_t>(ps) * p.rot + half + i];
              x[e] = lo ? x[e] * c - y[e] * s : x[e] * c + y[e] * s;
            }
          }
        } else {
          if (e0 < p.rot) {
#pragma unroll
            for (int e = 0; e < C; e += 2) {
              const int i = (e0 + e) >> 1;
              const int ps = pos[mrope_axis(i, p.sec0, p.sec1)];
              const float c = p.cos_sin[static_cast<size_t>(ps) * p.rot + i];
              const float s = p.cos_sin[static_cast<size_t>(ps) * p.rot + half + i];
              const float a = x[e], b = x[e + 1];
              x[e] = a * c - b * s;
              x[e + 1] = b * c + a * s;
            }
          }
        }
      }
      uint32_t packed[C / 2];
#pragma unroll
      for (int e = 0; e < C; e += 2) packed[e / 2] = pack_bf16(x[e], x[e + 1]);
#pragma unroll
      for (int e = 0; e < C / 2; ++e) *reinterpret_cast<uint32_t*>(ptr + e0 + 2 * e) = packed[e];
      if (is_k && slot >= 0 && p.k_cache != nullptr) {
        const int page = slot / p.page_size, off = slot - page * p.page_size;
        const size_t base = ((static_cast<size_t>(page) * p.Hkv + h) * (D / 64) + e0 / 64) * p.page_size * 64 +
                            static_cast<size_t>(off) * 64 + (e0 % 64);
#pragma unroll
        for (int e = 0; e < C / 2; ++e) *reinterpret_cast<uint32_t*>(p.k_cache + base + 2 * e) = packed[e];
      }
    } else if (slot >= 0) {
      const __nv_bfloat16* ptr = p.v + t * p.v_ts + h * p.v_hs;
      const int page = slot / p.page_size, off = slot - page * p.page_size;
      const size_t base = ((static_cast<size_t>(page) * p.Hkv + h) * (D / 64) + e0 / 64) * p.page_size * 64 +
                          static_cast<size_t>(off) * 64 + (e0 % 64);
#pragma unroll
      for (int e = 0; e < C; e += 2)
        *reinterpret_cast<uint32_t*>(p.v_cache + base + e) = *reinterpret_cast<const uint32_t*>(ptr + e0 + e);
    }
  }
}

}  // namespace b200

using namespace b200;

GLLM_EXPORT int gllm_rope_kv_write(void* q, int64_t q_ts, int64_t q_hs, int Hq, void* k, int64_t k_ts,
                                   int64_t k_hs, int Hkv, const void* v, int64_t v_ts, int64_t v_hs,
                                   const void* q_norm_w, const void* k_norm_w, const void* cos_sin, int D,
                                   int rot, int neox, int T, const void* positions, const void* slots,
                                   float eps, void* k_cache, void* v_cache, int sec0, int sec1,
                                   int page_size, int64_t pos_stride, void* stream) {
  if (T <= 0) return 0;
  RopeParams p;
  p.q = reinterpret_cast<__nv_bfloat16*>(q); p.q_ts = q_ts; p.q_hs = q_hs; p.Hq = Hq;
  p.k = reinterpret_cast<__nv_bfloat16*>(k); p.k_ts = k_ts; p.k_hs = k_hs; p.Hkv = Hkv;
  p.v = reinterpret_cast<const __nv_bfloat16*>(v); p.v_ts = v_ts; p.v_hs = v_hs;
  p.q_norm_w = reinterpret_cast<const __nv_bfloat16*>(q_norm_w);
  p.k_norm_w = reinterpret_cast<const __nv_bfloat16*>(k_norm_w);
  p.cos_sin = reinterpret_cast<const float*>(cos_sin);
  p.D = D; p.rot = rot; p.neox = neox; p.T = T;
  p.positions = reinterpret_cast<const int32_t*>(positions);
  p.slots = reinterpret_cast<const int32_t*>(slots);
  p.eps = eps;
  p.k_cache = reinterpret_cast<__nv_bfloat16*>(k_cache);
  p.v_cache = reinterpret_cast<__nv_bfloat16*>(v_cache);
  p.sec0 = sec0; p.sec1 = sec1; p.page_size = page_size; p.pos_stride = pos_stride;
  const int C = D / 32;
  if (D % 64 != 0 || (rot > 0 && ((rot / 2) % C != 0 || rot > D))) {
    fprintf(stderr, "[gllm_b200] rope_kv_write: unsupported D=%d rot=%d\n", D, rot);
    return 1;
  }
  const int n_heads = Hq + 2 * Hkv;
  int warps = n_heads < 4 ? n_heads : 4;
  if (warps < 1) warps = 1;
  const dim3 grid(T, (n_heads + warps - 1) / warps);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (C) {
    case 2: CUDA_CHECK_RET(launch_pdl(rope_kv_kernel<2>, grid, dim3(warps * 32), 0, st, p)); break;
    case 4: CUDA_CHECK_RET(launch_pdl(rope_kv_kernel<4>, grid, dim3(warps * 32), 0, st, p)); break;
    case 8: CUDA_CHECK_RET(launch_pdl(rope_kv_kernel<8>, grid, dim3(warps * 32), 0, st, p)); break;
    default:
      fprintf(stderr, "[gllm_b200] rope_kv_write: unsupported head_dim %d\n", D);
      return 1;
  }
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
