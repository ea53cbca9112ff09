// Micro-benchmark: how fast can SMs stream a large bf16 matrix from HBM into shared memory?
// Guides the weight-streaming design of the decode GEMM (profiles/tma_stream.md).
//   mode 0: TMA 2D boxes {64 cols, ROWS rows}, SWIZZLE_128B, row pitch = ld (strided 128 B segments)
//   mode 1: 1D bulk copies (cp.async.bulk.shared.global) of `chunk` contiguous bytes
//   mode 2: plain vectorised LDG (16 B / thread), no smem
// Each CTA streams its own contiguous share of the tiles through a `stages`-deep mbarrier ring.
#include "../common/host_utils.h"
#include "../common/ptx.cuh"
#include <string.h>

namespace b200 {

__global__ void tma_stream_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* base, int mode,
                                  int box_rows, int chunk_bytes, int tiles_total, int tiles_k, int stages,
                                  unsigned long long* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = mode == 0 ? box_rows * 128 : chunk_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  if (mode == 2) {
    // each thread sums 16-byte words
    const size_t total16 = static_cast<size_t>(tiles_total) * chunk_bytes / 16;
    unsigned long long acc = 0;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total16;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
      uint4 v = ld_nc_v4(base + i * 16);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x1234567ull) *sink = acc;
    return;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // single thread: issue `stages` loads ahead, then wait/reissue (no consumer work)
    int issued = 0, done = 0;
    const int my_first = blockIdx.x, stride = gridDim.x;
    auto issue = [&](int idx) {
      const int tile = my_first + idx * stride;
      const int s = idx % stages;
      mbar_expect_tx(&full[s], stage_bytes);
      if (mode == 0) {
        const int tk = tile % tiles_k, tr = tile / tiles_k;
        tma_load_2d(smem + s * stage_bytes, &tmap, &full[s], tk * 64, tr * box_rows, kEvictFirst);
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + s * stage_bytes)), "l"(base + static_cast<size_t>(tile) * chunk_bytes),
                       "r"(stage_bytes), "r"(smem_u32(&full[s])) : "memory");
      }
    };
    const int n_mine = (tiles_total - my_first + stride - 1) / stride;
    while (issued < n_mine && issued < stages) issue(issued++);
    while (done < n_mine) {
      mbar_wait(&full[done % stages], (done / stages) & 1);
      ++done;
      if (issued < n_mine) issue(issued++);
    }
  }
}

}  // namespace b200

using namespace b200;

GLLM_EXPORT int gllm_bench_tma_stream(const void* base, int64_t rows, int64_t cols, int64_t ld, int mode,
                                      int box_rows, int chunk_bytes, int stages, int ctas_per_sm, void* sink,
                                      void* stream) {
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  int tiles_total, tiles_k = 1;
  if (mode == 0) {
    if (make_tmap_2d(&tm, base, rows, cols, ld * 2, box_rows, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16)) return 1;
    tiles_k = static_cast<int>(cols / 64);
    tiles_total = static_cast<int>(rows / box_rows) * tiles_k;
  } else {
    tiles_total = static_cast<int>(rows * cols * 2 / chunk_bytes);
  }
  const int stage_bytes = mode == 0 ? box_rows * 128 : chunk_bytes;
  const int smem = mode == 2 ? 0 : stages * stage_bytes + 1024 + 256;
  CUDA_CHECK_RET(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int grid = num_sms() * ctas_per_sm;
  tma_stream_kernel<<<grid, mode == 2 ? 512 : 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      tm, reinterpret_cast<const uint8_t*>(base), mode, box_rows, chunk_bytes, tiles_total, tiles_k, stages,
      reinterpret_cast<unsigned long long*>(sink));
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}
