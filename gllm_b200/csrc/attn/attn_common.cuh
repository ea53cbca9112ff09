// Shared pieces of the paged attention kernels (GQA flash attention in paged_attention.cu, MLA latent attention
// in mla_attention.cu): TMA page-box loads, ldmatrix / mma.sync wrappers, swizzled tile addressing and the
// cached KV tensor maps.
#pragma once

#include <mutex>
#include <unordered_map>

#include "../common/host_utils.h"
#include "../common/ptx.cuh"

namespace b200 {

static constexpr int kTileN = 64;  // KV tokens per pipeline stage
static constexpr int kMathWarps = 4;
static constexpr int kAttnThreads = (kMathWarps + 1) * 32;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Byte offset of (token row `tok` in [0,64), 16-byte chunk `c8` in [0, D/8)) inside a K/V tile that
// was assembled from TMA page boxes [D/64][page_size][64] with the 128-byte swizzle.
__device__ __forceinline__ uint32_t tile_off(int tok, int c8, int page_size, int page_bytes) {
  const int pj = tok / page_size;
  const int r = tok - pj * page_size;
  return pj * page_bytes + (c8 >> 3) * (page_size * 128) + r * 128 + (((c8 & 7) ^ (r & 7)) << 4);
}

struct TmapKey {
  const void* base;
  int64_t pages;
  int hkv, d, page;
  bool operator==(const TmapKey& o) const {
    return base == o.base && pages == o.pages && hkv == o.hkv && d == o.d && page == o.page;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    return std::hash<const void*>()(k.base) ^ (std::hash<int64_t>()(k.pages) << 1) ^ (k.hkv * 1315423911u) ^
           (k.d * 2654435761u) ^ (k.page * 97u);
  }
};

inline int get_kv_tmap(const void* base, int64_t num_pages, int Hkv, int D, int page_size, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{base, num_pages, Hkv, D, page_size};
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  CUtensorMap m;
  cuuint64_t dims[3] = {64, (cuuint64_t)page_size, (cuuint64_t)(num_pages * Hkv * (D / 64))};
  cuuint64_t strides[2] = {128, (cuuint64_t)page_size * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)page_size, (cuuint32_t)(D / 64)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[gllm_b200] KV tensor-map encode failed (%d)\n", (int)r);
    return 1;
  }
  cache.emplace(key, m);
  *out = m;
  return 0;
}


}  // namespace b200
