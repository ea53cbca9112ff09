// Host-side helpers shared by all launchers: error checks, TMA tensor-map
// encoding (driver entry point resolved at runtime, so the library never links
// libcuda directly and can be built on a GPU-less box), device properties.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define GLLM_EXPORT extern "C" __attribute__((visibility("default")))

#define CUDA_CHECK_RET(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      fprintf(stderr, "[gllm_b200] CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__, \
              cudaGetErrorString(_e));                                                    \
      return static_cast<int>(_e);                                                        \
    }                                                                                     \
  } while (0)

namespace b200 {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr) {
      fprintf(stderr, "[gllm_b200] cannot resolve cuTensorMapEncodeTiled (%s)\n",
              cudaGetErrorString(e));
      abort();
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D row-major tensor [rows, cols] of `elem_bytes`-sized elements with row pitch
// `ld_bytes`; box = [box_rows, box_cols]; 128-byte swizzle (box_cols*elem_bytes == 128)
// unless `swizzle128` is false.
inline int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t ld_bytes, uint32_t box_rows, uint32_t box_cols,
                        CUtensorMapDataType dtype, bool swizzle128 = true) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(
      map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr,
            "[gllm_b200] cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu "
            "box=%ux%u\n",
            (int)r, base, (unsigned long long)rows, (unsigned long long)cols,
            (unsigned long long)ld_bytes, box_rows, box_cols);
    return 1;
  }
  return 0;
}

// 3-D tensor [d2][d1][d0] (d0 contiguous) with byte strides for d1 and d2; box = [1][box_d1][box_d0], 128-byte
// swizzle (box_d0 * elem_bytes == 128). Used for batched / strided operands (a [T, heads, K] activation read as
// per-head [T, K] matrices without a transposing copy).
inline int make_tmap_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box_d0, uint32_t box_d1,
                        CUtensorMapDataType dtype) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box_d0, box_d1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[gllm_b200] cuTensorMapEncodeTiled(3d) failed (%d): base=%p dims=%llu,%llu,%llu strides=%llu,%llu\n",
            (int)r, base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
            (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
    return 1;
  }
  return 0;
}

// Programmatic dependent launch: the kernel may start (prologue, weight prefetch) while its predecessor in the
// stream is still draining; it must execute `griddep_wait()` (ptx.cuh) before touching anything the
// predecessor wrote. Only kernels written that way are launched through this helper. GLLM_PDL=0 disables it.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GLLM_PDL");
    v = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// `cudaFuncSetAttribute` is per DEVICE: a launcher remembers, per kernel instantiation, on which devices the
// dynamic shared memory opt-in has been done (one process normally drives one GPU, but tests and tools may not).
struct PerDeviceOnce {
  unsigned long long mask = 0;   // devices 0..63
  static unsigned long long bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
  }
  bool need() const { return (__atomic_load_n(&mask, __ATOMIC_ACQUIRE) & bit()) == 0; }
  void done() { __atomic_fetch_or(&mask, bit(), __ATOMIC_RELEASE); }
};

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace b200
