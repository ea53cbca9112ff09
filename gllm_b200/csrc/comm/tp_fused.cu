// Tensor-parallel collectives fused into the compute kernels, over NVLink peer memory (P2P ld/st
// on symmetric buffers). Together with the hooks in gemm/gemm_bf16.cu these implement the
// token-sharded ("sequence parallel") block dataflow that replaces the reference's
// GEMM -> NCCL all_reduce -> fused_add_rms_norm (gllm/layers/linear.py:247-250):
//
//   row-parallel GEMM  ⊕ reduce-scatter : the GEMM epilogue stores every partial tile straight into
//        the owner rank's staging slot   stage[owner][src][row_local][H]   and bumps
//        cnt[owner][src] with red.release.sys   (gemm_bf16.cu, rs_* params)
//   rs_reduce_norm (this file)          : owner waits for the tile counters, sums the tp partials in a
//        fixed order (deterministic), adds the residual shard, applies RMSNorm and PUSHES the normed
//        rows into every peer's gather buffer (all-gather by producer-side stores), then publishes an
//        epoch flag per peer with st.release.sys
//   column-parallel GEMM ⊕ all-gather   : the next GEMM's TMA producer warp spins (ld.acquire.sys) on
//        the flags of the row shards that cover its M tile before loading A   (gemm_bf16.cu, a_ready)
//
// Everything is CUDA-graph safe: no host-side epochs — expected counter values and flag epochs live
// in device memory and are advanced by the kernels themselves.
#include "../common/host_utils.h"
#include "../common/ptx.cuh"

namespace b200 {

static constexpr int kMaxTp = 8;
static constexpr int kMaxBlocks = 256;  // 128-row blocks per gather buffer (32 K tokens)
static constexpr int kFlagBlockRows = 128;

struct TpState {
  // device-resident, one per rank (NOT symmetric): advanced by the kernels
  uint32_t rs_expected[2][kMaxTp];  // per parity, per source rank: tile arrivals consumed so far
  uint32_t ag_epoch[3];             // per gather buffer: number of pushes so far (diagnostic)
  uint32_t ticket[4];               // grid tickets
  uint32_t ll_epoch;                // one-shot LL all-reduce: calls completed so far (flag value of the next call - 1)
  uint32_t nvls_epoch;              // NVLS (multimem) all-reduce: calls completed so far
  uint32_t pad[7];                  // -> ag_expected starts at byte 128
  // per gather buffer, per 128-row block: rows that must have arrived before the block may be read.
  // Arrival counters (symmetric `flags`) are bumped once per pushed row by the row's owner, so a
  // consumer GEMM can start on the blocks that are complete while the rest is still in flight.
  uint32_t ag_expected[3][kMaxBlocks];
};

struct ReduceNormParams {
  // partial sources: stage slots of this rank (local memory), [src][rows_per_rank][H]
  const __nv_bfloat16* stage;
  const uint32_t* cnt;          // [kMaxTp] arrival counters for this parity (local, written by peers)
  uint32_t n_tiles[kMaxTp];     // arrivals expected from each source this call
  const __nv_bfloat16* local_x; // != null: single local source (rows of this rank's shard), no waiting
  int64_t local_ld;
  __nv_bfloat16* residual;      // [rows_per_rank, H] shard (in/out); may be null on first use
  int residual_in;              // 0: residual := sum (first layer), 1: residual += sum
  const __nv_bfloat16* norm_w;
  __nv_bfloat16* ag_peers[kMaxTp];  // gather buffer [T_pad, H] of every rank (peer pointers)
  uint32_t* flag_peers[kMaxTp];     // flags[tp] of every rank for this gather buffer
  __nv_bfloat16* unnormed_out;      // optional: also store the un-normalised sum (PP boundary)
  TpState* st;
  int parity, ag_idx;
  int tp, rank, rows_per_rank, rows_valid, H;
  float eps;
  int T;  // total tokens of this forward (all ranks)
  int bcast;  // one-shot all-reduce mode: every rank holds and reduces ALL T rows (rows_per_rank == T)
  // bcast + push_x: the kernel first publishes this rank's partial row to every rank's staging slot (one-shot
  // all-reduce ⊕ add ⊕ RMSNorm in ONE kernel for decode-sized T)
  const __nv_bfloat16* push_x;
  int64_t push_ld;
  __nv_bfloat16* stage_peers[kMaxTp];
  uint32_t* cnt_peers[kMaxTp];
};

template <int NV>
__global__ void rs_reduce_norm_kernel(const ReduceNormParams p) {
  __shared__ float red[32];
  __shared__ uint32_t s_last;
  const int row = blockIdx.x;  // local row in this rank's shard
  const int nvec = p.H >> 3;

  if (p.push_x != nullptr) {
    if (row < p.rows_valid) {
      const __nv_bfloat16* src = p.push_x + static_cast<size_t>(row) * p.push_ld;
      for (int i = threadIdx.x * 8; i < p.H; i += blockDim.x * 8) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i);
        for (int d = 0; d < p.tp; ++d) {
          const int peer = (p.rank + d) % p.tp;
          st_v4(p.stage_peers[peer] + (static_cast<size_t>(p.rank) * p.rows_per_rank + row) * p.H + i, v);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0 && row < p.rows_valid) {
      __threadfence_system();
      for (int d = 0; d < p.tp; ++d) red_add_relaxed_sys(p.cnt_peers[(p.rank + d) % p.tp] + p.rank, 1u);
    }
  }
  if (p.local_x == nullptr && threadIdx.x < p.tp) {
    const int src = threadIdx.x;
    const uint32_t target = p.st->rs_expected[p.parity][src] + p.n_tiles[src];
    const uint32_t* c = p.cnt + src;
    // wrap-safe comparison on monotonically increasing counters
    SpinGuard guard;
    while (static_cast<int32_t>(ld_acquire_sys(c) - target) < 0) guard.poll();
  }
  __syncthreads();

  float v[NV][8];
  float ss = 0.f;
  if (row < p.rows_valid) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * blockDim.x;
      if (i < nvec) {
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.local_x != nullptr) {
          const uint4 t = *reinterpret_cast<const uint4*>(
              p.local_x + (p.bcast ? static_cast<size_t>(row) : static_cast<size_t>(p.rank) * p.rows_per_rank + row) * p.local_ld + i * 8);
          const float2 f0 = unpack_bf16(t.x), f1 = unpack_bf16(t.y), f2 = unpack_bf16(t.z), f3 = unpack_bf16(t.w);
          a[0] = f0.x; a[1] = f0.y; a[2] = f1.x; a[3] = f1.y; a[4] = f2.x; a[5] = f2.y; a[6] = f3.x; a[7] = f3.y;
        } else {
          for (int s = 0; s < p.tp; ++s) {
            const uint4 t = __ldcg(reinterpret_cast<const uint4*>(
                p.stage + (static_cast<size_t>(s) * p.rows_per_rank + row) * p.H + i * 8));
            const float2 f0 = unpack_bf16(t.x), f1 = unpack_bf16(t.y), f2 = unpack_bf16(t.z), f3 = unpack_bf16(t.w);
            a[0] += f0.x; a[1] += f0.y; a[2] += f1.x; a[3] += f1.y; a[4] += f2.x; a[5] += f2.y; a[6] += f3.x; a[7] += f3.y;
          }
        }
        if (p.unnormed_out != nullptr) {
          uint4 o;
          o.x = pack_bf16(a[0], a[1]); o.y = pack_bf16(a[2], a[3]); o.z = pack_bf16(a[4], a[5]); o.w = pack_bf16(a[6], a[7]);
          *reinterpret_cast<uint4*>(p.unnormed_out + static_cast<size_t>(row) * p.H + i * 8) = o;
        }
        if (p.residual != nullptr) {
          __nv_bfloat16* rp = p.residual + static_cast<size_t>(row) * p.H + i * 8;
          if (p.residual_in) {
            const uint4 r = *reinterpret_cast<const uint4*>(rp);
            const float2 r0 = unpack_bf16(r.x), r1 = unpack_bf16(r.y), r2 = unpack_bf16(r.z), r3 = unpack_bf16(r.w);
            // the all-reduced value is rounded to bf16 before the add, like the unfused reference
            a[0] = __bfloat162float(__float2bfloat16(a[0])) + r0.x; a[1] = __bfloat162float(__float2bfloat16(a[1])) + r0.y;
            a[2] = __bfloat162float(__float2bfloat16(a[2])) + r1.x; a[3] = __bfloat162float(__float2bfloat16(a[3])) + r1.y;
            a[4] = __bfloat162float(__float2bfloat16(a[4])) + r2.x; a[5] = __bfloat162float(__float2bfloat16(a[5])) + r2.y;
            a[6] = __bfloat162float(__float2bfloat16(a[6])) + r3.x; a[7] = __bfloat162float(__float2bfloat16(a[7])) + r3.y;
          }
          uint4 o;
          o.x = pack_bf16(a[0], a[1]); o.y = pack_bf16(a[2], a[3]); o.z = pack_bf16(a[4], a[5]); o.w = pack_bf16(a[6], a[7]);
          *reinterpret_cast<uint4*>(rp) = o;
          const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y), q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
          a[0] = q0.x; a[1] = q0.y; a[2] = q1.x; a[3] = q1.y; a[4] = q2.x; a[5] = q2.y; a[6] = q3.x; a[7] = q3.y;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[j][e] = a[e]; ss += a[e] * a[e]; }
      }
    }
  }
  // block reduce
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float t = lane < ((blockDim.x + 31) >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    ss = t;
  }
  if (row < p.rows_valid && p.norm_w != nullptr) {
    const float inv = rsqrtf(ss / static_cast<float>(p.H) + p.eps);
    const size_t grow = p.bcast ? static_cast<size_t>(row) : static_cast<size_t>(p.rank) * p.rows_per_rank + row;
    const int n_push = p.bcast ? 1 : p.tp;  // bcast: the result stays local (every rank computes all rows)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int i = threadIdx.x + j * blockDim.x;
      if (i < nvec) {
        const uint4 wv = *reinterpret_cast<const uint4*>(p.norm_w + i * 8);
        const float2 w0 = unpack_bf16(wv.x), w1 = unpack_bf16(wv.y), w2 = unpack_bf16(wv.z), w3 = unpack_bf16(wv.w);
        uint4 o;
        o.x = pack_bf16(v[j][0] * inv * w0.x, v[j][1] * inv * w0.y);
        o.y = pack_bf16(v[j][2] * inv * w1.x, v[j][3] * inv * w1.y);
        o.z = pack_bf16(v[j][4] * inv * w2.x, v[j][5] * inv * w2.y);
        o.w = pack_bf16(v[j][6] * inv * w3.x, v[j][7] * inv * w3.y);
        // all-gather by producer-side stores: own copy first, then the peers (rank-rotated order)
        for (int d = 0; d < n_push; ++d) {
          const int peer = (p.rank + d) % p.tp;
          st_v4(p.ag_peers[peer] + grow * p.H + i * 8, o);
        }
      }
    }
  }
  // publish this row: one arrival on the row's 128-row block counter of every rank
  // (CTA barrier, then ONE system fence by the signalling thread: cumulative over the CTA's stores)
  __syncthreads();
  if (threadIdx.x == 0) {
    if (row < p.rows_valid && p.norm_w != nullptr) {
      __threadfence_system();
      const int blk = (p.bcast ? row : p.rank * p.rows_per_rank + row) / kFlagBlockRows;
      for (int d = 0; d < (p.bcast ? 1 : p.tp); ++d) {
        const int peer = (p.rank + d) % p.tp;
        red_add_relaxed_sys(p.flag_peers[peer] + blk, 1u);
      }
    }
    const uint32_t old = atomicAdd(&p.st->ticket[0], 1u);
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  // the last CTA of the grid advances the device-side bookkeeping (no host epochs => graph safe)
  if (s_last) {
    if (threadIdx.x == 0) {
      p.st->ticket[0] = 0u;
      if (p.local_x == nullptr)
        for (int s = 0; s < p.tp; ++s) p.st->rs_expected[p.parity][s] += p.n_tiles[s];
      p.st->ag_epoch[p.ag_idx] += 1u;
    }
    // every row < T is pushed by its owner: block b will receive min(T, (b+1)*128) - b*128 arrivals
    const int nblk = (p.T + kFlagBlockRows - 1) / kFlagBlockRows;
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
      const int rows = min(p.T, (b + 1) * kFlagBlockRows) - b * kFlagBlockRows;
      p.st->ag_expected[p.ag_idx][b] += static_cast<uint32_t>(rows);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One-shot all-reduce ⊕ residual add ⊕ RMSNorm for decode-sized T, low-latency ("LL") protocol: every 4 bytes
// of payload travel in an 8-byte store together with the call's epoch, so the receiver needs no fence, no
// counter and no second round trip — it polls the slot until the epoch matches. One CTA per token row; sums
// are taken in rank order (deterministic). Slots are double-buffered by call parity (a peer is at most one
// call ahead). Replaces NCCL all_reduce + fused_add_rms_norm (reference: gllm/layers/linear.py:247-250).
// ---------------------------------------------------------------------------------------------
struct LLParams {
  const __nv_bfloat16* x;      // this rank's partial [T, H]
  int64_t ldx;
  __nv_bfloat16* residual;     // [T, H] replicated running residual (in/out)
  int residual_in;
  const __nv_bfloat16* norm_w;
  __nv_bfloat16* out;          // [T, H] normed output (local)
  uint2* ll_peers[kMaxTp];     // LL slot base of this parity on every rank: [src][row_cap][H/2] x 8 bytes
  TpState* st;
  int tp, rank, T, H, row_cap;
  float eps;
};

__device__ __forceinline__ void st_volatile_v2(uint2* p, uint32_t a, uint32_t b) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void ll_allreduce_norm_kernel(const LLParams p) {
  __shared__ float red[32];
  __shared__ uint32_t s_last;
  griddep_launch();  // the consumer GEMM may start streaming its weights while we exchange partials
  griddep_wait();
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const bool active = tid * 8 < p.H;
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(&p.st->ll_epoch) + 1u;
  const size_t slot0 = static_cast<size_t>(row) * (p.H / 2) + tid * 4;
  const size_t src_stride = static_cast<size_t>(p.row_cap) * (p.H / 2);
  uint32_t mine[4] = {0, 0, 0, 0};
  if (active) {
    const uint4 v = *reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(row) * p.ldx + tid * 8);
    mine[0] = v.x; mine[1] = v.y; mine[2] = v.z; mine[3] = v.w;
    for (int d = 1; d < p.tp; ++d) {
      uint2* dst = p.ll_peers[(p.rank + d) % p.tp] + p.rank * src_stride + slot0;
#pragma unroll
      for (int k = 0; k < 4; ++k) st_volatile_v2(dst + k, mine[k], epoch);
    }
  }
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    SpinGuard guard;
    const uint2* mybuf = p.ll_peers[p.rank];
    for (int s = 0; s < p.tp; ++s) {
      uint32_t w[4];
      if (s == p.rank) {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = mine[k];
      } else {
        const uint2* src = mybuf + s * src_stride + slot0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint2 v;
          v = ld_volatile_v2(src + k);
          while (v.y != epoch) {
            guard.poll();
            v = ld_volatile_v2(src + k);
          }
          w[k] = v.x;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16(w[k]);
        a[2 * k] += f.x; a[2 * k + 1] += f.y;
      }
    }
  }
  float ss = 0.f;
  if (active) {
    __nv_bfloat16* rp = p.residual + static_cast<size_t>(row) * p.H + tid * 8;
    if (p.residual_in) {
      const uint4 r = *reinterpret_cast<const uint4*>(rp);
      const float2 r0 = unpack_bf16(r.x), r1 = unpack_bf16(r.y), r2 = unpack_bf16(r.z), r3 = unpack_bf16(r.w);
      // the all-reduced value is rounded to bf16 before the add, like the unfused reference
      a[0] = __bfloat162float(__float2bfloat16(a[0])) + r0.x; a[1] = __bfloat162float(__float2bfloat16(a[1])) + r0.y;
      a[2] = __bfloat162float(__float2bfloat16(a[2])) + r1.x; a[3] = __bfloat162float(__float2bfloat16(a[3])) + r1.y;
      a[4] = __bfloat162float(__float2bfloat16(a[4])) + r2.x; a[5] = __bfloat162float(__float2bfloat16(a[5])) + r2.y;
      a[6] = __bfloat162float(__float2bfloat16(a[6])) + r3.x; a[7] = __bfloat162float(__float2bfloat16(a[7])) + r3.y;
    }
    uint4 o;
    o.x = pack_bf16(a[0], a[1]); o.y = pack_bf16(a[2], a[3]); o.z = pack_bf16(a[4], a[5]); o.w = pack_bf16(a[6], a[7]);
    *reinterpret_cast<uint4*>(rp) = o;
    const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y), q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
    a[0] = q0.x; a[1] = q0.y; a[2] = q1.x; a[3] = q1.y; a[4] = q2.x; a[5] = q2.y; a[6] = q3.x; a[7] = q3.y;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += a[e] * a[e];
  }
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float t = lane < ((blockDim.x + 31) >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    ss = t;
  }
  if (active) {
    const float inv = rsqrtf(ss / static_cast<float>(p.H) + p.eps);
    const uint4 wv = *reinterpret_cast<const uint4*>(p.norm_w + tid * 8);
    const float2 w0 = unpack_bf16(wv.x), w1 = unpack_bf16(wv.y), w2 = unpack_bf16(wv.z), w3 = unpack_bf16(wv.w);
    uint4 o;
    o.x = pack_bf16(a[0] * inv * w0.x, a[1] * inv * w0.y);
    o.y = pack_bf16(a[2] * inv * w1.x, a[3] * inv * w1.y);
    o.z = pack_bf16(a[4] * inv * w2.x, a[5] * inv * w2.y);
    o.w = pack_bf16(a[6] * inv * w3.x, a[7] * inv * w3.y);
    *reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row) * p.H + tid * 8) = o;
  }
  // the last CTA to finish closes the epoch (all CTAs are co-resident: T <= row_cap <= 64)
  __syncthreads();
  if (tid == 0) {
    const uint32_t old = atomicAdd(&p.st->ticket[1], 1u);
    s_last = (old == gridDim.x - 1) ? 1u : 0u;
    if (s_last) {
      p.st->ticket[1] = 0u;
      p.st->ll_epoch = epoch;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// NVLS one-shot all-reduce ⊕ residual add ⊕ RMSNorm: the reduction happens INSIDE the NVSwitch. Every rank stores its
// partial row into its own slice of a symmetric buffer that is also mapped as a multicast object, publishes
// "row r of rank s is written" with ONE multimem.red on a multicast flag (lands in every rank's flag array), waits
// until all tp sources have published the row, and then reads the row through the multicast address with
// multimem.ld_reduce: the switch fetches the tp replicas and returns their fp32-accumulated sum. Per rank and row
// that is H*2 bytes in and H*2 bytes out over NVLink, independent of tp — the LL variant above receives
// (tp-1) * H * 4 bytes. Replaces NCCL all_reduce + add + norm of the reference (gllm/layers/linear.py:247-250,
// gllm/layers/layernorm.py) for decode-sized batches when the fabric supports multicast.
// ------------------------------------------------------------------------------------------------------------------
struct NvlsParams {
  const __nv_bfloat16* x;      // this rank's partial [T, H]
  int64_t ldx;
  __nv_bfloat16* residual;     // [T, H] replicated running residual (in/out)
  int residual_in;
  const __nv_bfloat16* norm_w;
  __nv_bfloat16* out;          // [T, H] normed output (local)
  __nv_bfloat16* buf_local;    // this rank's [row_cap, H] slice for this parity (unicast address)
  const __nv_bfloat16* buf_mc; // the same slice through the multicast mapping
  uint32_t* flags_mc;          // [tp][row_cap] epoch flags, multicast address (a store reaches every rank)
  const uint32_t* flags_local; // this rank's copy of the flags (unicast address)
  TpState* st;
  int tp, rank, T, H, row_cap;
  float eps;
};

__device__ __forceinline__ void multimem_red_max_release(uint32_t* mc_addr, uint32_t v) {
  asm volatile("multimem.red.release.sys.global.max.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
// 8 bf16 (16 bytes) of the sum over all replicas, accumulated in fp32 inside the switch
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_addr) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}

__global__ void nvls_allreduce_norm_kernel(const NvlsParams p) {
  __shared__ float red[32];
  griddep_launch();
  griddep_wait();
  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const bool active = tid * 8 < p.H;
  const uint32_t epoch = *reinterpret_cast<const volatile uint32_t*>(&p.st->nvls_epoch) + 1u;
  const size_t off = static_cast<size_t>(row) * p.H + tid * 8;
  if (active) {
    const uint4 v = *reinterpret_cast<const uint4*>(p.x + static_cast<size_t>(row) * p.ldx + tid * 8);
    *reinterpret_cast<uint4*>(p.buf_local + off) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    // one instruction tells every rank that this rank's row is in place (max: the flag only ever moves forward)
    multimem_red_max_release(p.flags_mc + p.rank * p.row_cap + row, epoch);
    SpinGuard guard;
    for (int s = 0; s < p.tp; ++s) {
      while (static_cast<int32_t>(ld_acquire_sys(p.flags_local + s * p.row_cap + row) - epoch) < 0) guard.poll();
    }
  }
  __syncthreads();
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (active) {
    const uint4 v = multimem_ld_reduce_bf16x8(p.buf_mc + off);
    const float2 f0 = unpack_bf16(v.x), f1 = unpack_bf16(v.y), f2 = unpack_bf16(v.z), f3 = unpack_bf16(v.w);
    a[0] = f0.x; a[1] = f0.y; a[2] = f1.x; a[3] = f1.y; a[4] = f2.x; a[5] = f2.y; a[6] = f3.x; a[7] = f3.y;
  }
  float ss = 0.f;
  if (active) {
    __nv_bfloat16* rp = p.residual + off;
    if (p.residual_in) {
      const uint4 r = *reinterpret_cast<const uint4*>(rp);
      const float2 r0 = unpack_bf16(r.x), r1 = unpack_bf16(r.y), r2 = unpack_bf16(r.z), r3 = unpack_bf16(r.w);
      a[0] += r0.x; a[1] += r0.y; a[2] += r1.x; a[3] += r1.y; a[4] += r2.x; a[5] += r2.y; a[6] += r3.x; a[7] += r3.y;
    }
    uint4 o;
    o.x = pack_bf16(a[0], a[1]); o.y = pack_bf16(a[2], a[3]); o.z = pack_bf16(a[4], a[5]); o.w = pack_bf16(a[6], a[7]);
    *reinterpret_cast<uint4*>(rp) = o;
    const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y), q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
    a[0] = q0.x; a[1] = q0.y; a[2] = q1.x; a[3] = q1.y; a[4] = q2.x; a[5] = q2.y; a[6] = q3.x; a[7] = q3.y;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss += a[e] * a[e];
  }
  {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float t = lane < ((blockDim.x + 31) >> 5) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    ss = t;
  }
  if (active) {
    const float inv = rsqrtf(ss / static_cast<float>(p.H) + p.eps);
    const uint4 wv = *reinterpret_cast<const uint4*>(p.norm_w + tid * 8);
    const float2 w0 = unpack_bf16(wv.x), w1 = unpack_bf16(wv.y), w2 = unpack_bf16(wv.z), w3 = unpack_bf16(wv.w);
    uint4 o;
    o.x = pack_bf16(a[0] * inv * w0.x, a[1] * inv * w0.y);
    o.y = pack_bf16(a[2] * inv * w1.x, a[3] * inv * w1.y);
    o.z = pack_bf16(a[4] * inv * w2.x, a[5] * inv * w2.y);
    o.w = pack_bf16(a[6] * inv * w3.x, a[7] * inv * w3.y);
    *reinterpret_cast<uint4*>(p.out + off) = o;
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t old = atomicAdd(&p.st->ticket[2], 1u);
    if (old == gridDim.x - 1) {       // last CTA closes the epoch
      p.st->ticket[2] = 0u;
      p.st->nvls_epoch = epoch;
    }
  }
}

// Push a full-length partial [T, H] (e.g. the MoE block output, or anything not produced by the fused
// GEMM epilogue) into the owners' staging slots; one CTA per row, one arrival per row.
__global__ void push_partial_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int T, int H, int rank,
                                         int rows_per_rank, __nv_bfloat16* const* stage_peers,
                                         uint32_t* const* cnt_peers) {
  const int row = blockIdx.x;
  if (row >= T) return;
  const int owner = row / rows_per_rank;
  const int rl = row - owner * rows_per_rank;
  __nv_bfloat16* dst = stage_peers[owner] + (static_cast<size_t>(rank) * rows_per_rank + rl) * H;
  const __nv_bfloat16* src = x + static_cast<size_t>(row) * ldx;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) st_v4(dst + i, *reinterpret_cast<const uint4*>(src + i));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    red_add_relaxed_sys(cnt_peers[owner] + rank, 1u);
  }
}

// Block the stream until the shards of a gather buffer have been published for the current epoch
// (for consumers that are not the flag-aware GEMM: row gathers, router, ...).
__global__ void wait_ag_flags_kernel(const uint32_t* flags, const TpState* st, int ag_idx, int nblk) {
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
    const uint32_t e = st->ag_expected[ag_idx][b];
    SpinGuard guard;
    while (static_cast<int32_t>(ld_acquire_sys(flags + b) - e) < 0) guard.poll();
  }
}

}  // namespace b200

using namespace b200;

struct ReduceNormArgs {
  const void* stage;
  const void* cnt;
  uint32_t n_tiles[kMaxTp];
  const void* local_x;
  int64_t local_ld;
  void* residual;
  int residual_in;
  const void* norm_w;
  void* ag_peers[kMaxTp];
  void* flag_peers[kMaxTp];
  void* unnormed_out;
  void* st;
  int parity, ag_idx, tp, rank, rows_per_rank, rows_valid, H;
  float eps;
  int T;
  int bcast;
  const void* push_x;
  int64_t push_ld;
  void* stage_peers[kMaxTp];
  void* cnt_peers[kMaxTp];
};

GLLM_EXPORT int gllm_rs_reduce_norm(const ReduceNormArgs* a, void* stream) {
  ReduceNormParams p;
  p.stage = reinterpret_cast<const __nv_bfloat16*>(a->stage);
  p.cnt = reinterpret_cast<const uint32_t*>(a->cnt);
  for (int i = 0; i < kMaxTp; ++i) {
    p.n_tiles[i] = a->n_tiles[i];
    p.ag_peers[i] = reinterpret_cast<__nv_bfloat16*>(a->ag_peers[i]);
    p.flag_peers[i] = reinterpret_cast<uint32_t*>(a->flag_peers[i]);
  }
  p.local_x = reinterpret_cast<const __nv_bfloat16*>(a->local_x);
  p.local_ld = a->local_ld;
  p.residual = reinterpret_cast<__nv_bfloat16*>(a->residual);
  p.residual_in = a->residual_in;
  p.norm_w = reinterpret_cast<const __nv_bfloat16*>(a->norm_w);
  p.unnormed_out = reinterpret_cast<__nv_bfloat16*>(a->unnormed_out);
  p.st = reinterpret_cast<TpState*>(a->st);
  p.parity = a->parity; p.ag_idx = a->ag_idx; p.tp = a->tp; p.rank = a->rank;
  p.rows_per_rank = a->rows_per_rank; p.rows_valid = a->rows_valid; p.H = a->H; p.eps = a->eps;
  p.T = a->T;
  p.bcast = a->bcast;
  p.push_x = reinterpret_cast<const __nv_bfloat16*>(a->push_x);
  p.push_ld = a->push_ld;
  for (int i = 0; i < kMaxTp; ++i) {
    p.stage_peers[i] = reinterpret_cast<__nv_bfloat16*>(a->stage_peers[i]);
    p.cnt_peers[i] = reinterpret_cast<uint32_t*>(a->cnt_peers[i]);
  }
  if ((p.T + kFlagBlockRows - 1) / kFlagBlockRows > kMaxBlocks) return 1;
  if (p.H % 8 != 0 || p.tp > kMaxTp) return 1;
  const int nvec = p.H / 8;
  int threads = ((nvec + 31) / 32) * 32, nv = 1;
  while (threads > 1024) { nv *= 2; threads = (((nvec + nv - 1) / nv + 31) / 32) * 32; }
  if (threads < 32) threads = 32;
  // one CTA per shard row; at least one CTA so the bookkeeping/flags advance even for empty shards
  const int grid = p.rows_per_rank > 0 ? p.rows_per_rank : 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // (plain launch: measured slower with programmatic dependent launch — the gated consumer GEMM then polls
  // system-scope flags next to the reduction, profiles/tp_fused.md)
  if (nv == 1) rs_reduce_norm_kernel<1><<<grid, threads, 0, st>>>(p);
  else if (nv == 2) rs_reduce_norm_kernel<2><<<grid, threads, 0, st>>>(p);
  else if (nv == 4) rs_reduce_norm_kernel<4><<<grid, threads, 0, st>>>(p);
  else return 1;
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

GLLM_EXPORT int gllm_push_partial_rows(const void* x, int64_t ldx, int T, int H, int rank, int rows_per_rank,
                                       const void* stage_peers_dev, const void* cnt_peers_dev, void* stream) {
  if (T <= 0) return 0;
  push_partial_rows_kernel<<<T, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, T, H, rank, rows_per_rank,
      reinterpret_cast<__nv_bfloat16* const*>(stage_peers_dev), reinterpret_cast<uint32_t* const*>(cnt_peers_dev));
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

// T: tokens of the current forward -> ceil(T / 128) row blocks to wait for
GLLM_EXPORT int gllm_wait_ag_flags(const void* flags, const void* st, int ag_idx, int T, void* stream) {
  const int nblk = (T + kFlagBlockRows - 1) / kFlagBlockRows;
  wait_ag_flags_kernel<<<1, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint32_t*>(flags), reinterpret_cast<const TpState*>(st), ag_idx, nblk);
  CUDA_CHECK_RET(cudaGetLastError());
  return 0;
}

struct LLArgs {
  const void* x;
  int64_t ldx;
  void* residual;
  int residual_in;
  const void* norm_w;
  void* out;
  void* ll_peers[kMaxTp];
  void* st;
  int tp, rank, T, H, row_cap;
  float eps;
};

GLLM_EXPORT int gllm_ll_allreduce_norm(const LLArgs* a, void* stream) {
  if (a->T <= 0) return 0;
  if (a->H % 8 != 0 || a->H / 8 > 1024 || a->T > a->row_cap || a->tp > kMaxTp) return 1;
  LLParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  p.ldx = a->ldx;
  p.residual = reinterpret_cast<__nv_bfloat16*>(a->residual);
  p.residual_in = a->residual_in;
  p.norm_w = reinterpret_cast<const __nv_bfloat16*>(a->norm_w);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  for (int i = 0; i < kMaxTp; ++i) p.ll_peers[i] = reinterpret_cast<uint2*>(a->ll_peers[i]);
  p.st = reinterpret_cast<TpState*>(a->st);
  p.tp = a->tp; p.rank = a->rank; p.T = a->T; p.H = a->H; p.row_cap = a->row_cap; p.eps = a->eps;
  const int threads = ((a->H / 8 + 31) / 32) * 32;
  CUDA_CHECK_RET(launch_pdl(ll_allreduce_norm_kernel, dim3(a->T), dim3(threads), 0,
                            reinterpret_cast<cudaStream_t>(stream), p));
  return 0;
}

struct NvlsArgs {
  const void* x;
  int64_t ldx;
  void* residual;
  int residual_in;
  const void* norm_w;
  void* out;
  void* buf_local;
  const void* buf_mc;
  void* flags_mc;
  const void* flags_local;
  void* st;
  int tp, rank, T, H, row_cap;
  float eps;
};

GLLM_EXPORT int gllm_nvls_allreduce_norm(const NvlsArgs* a, void* stream) {
  if (a->T <= 0) return 0;
  if (a->H % 8 != 0 || a->H / 8 > 1024 || a->T > a->row_cap || a->tp > kMaxTp) return 1;
  NvlsParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  p.ldx = a->ldx;
  p.residual = reinterpret_cast<__nv_bfloat16*>(a->residual);
  p.residual_in = a->residual_in;
  p.norm_w = reinterpret_cast<const __nv_bfloat16*>(a->norm_w);
  p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  p.buf_local = reinterpret_cast<__nv_bfloat16*>(a->buf_local);
  p.buf_mc = reinterpret_cast<const __nv_bfloat16*>(a->buf_mc);
  p.flags_mc = reinterpret_cast<uint32_t*>(a->flags_mc);
  p.flags_local = reinterpret_cast<const uint32_t*>(a->flags_local);
  p.st = reinterpret_cast<TpState*>(a->st);
  p.tp = a->tp; p.rank = a->rank; p.T = a->T; p.H = a->H; p.row_cap = a->row_cap; p.eps = a->eps;
  const int threads = ((a->H / 8 + 31) / 32) * 32;
  CUDA_CHECK_RET(launch_pdl(nvls_allreduce_norm_kernel, dim3(a->T), dim3(threads), 0,
                            reinterpret_cast<cudaStream_t>(stream), p));
  return 0;
}

GLLM_EXPORT int gllm_tp_state_bytes() { return static_cast<int>(sizeof(TpState)); }
