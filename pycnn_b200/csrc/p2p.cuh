// p2p.cuh -- data-parallel gradient exchange fused with the optimizer over NVLink peer memory.
//
// With one process per GPU the reference-equivalent step is  all-reduce(grads) -> ||g|| per tensor -> clip ->
// SGD/Adam on every rank.  NCCL does the first part in ~50-75 us for the 4.6 MB buffer of BASELINE configs[1]
// (latency-bound), and every rank then repeats the same norm + update over the whole buffer.  Here the three
// steps become one pipeline over peer pointers (CUDA IPC handles of every rank's grad/param/sync buffers):
//
//   A  p2p_reduce_kernel   rank r owns a contiguous range of the flat buffer's chunks; it waits for "grads
//                          ready" flags from all ranks, sums its range over all ranks' gradient buffers with
//                          128-bit loads across NVLink (fixed rank order: deterministic), stores the reduced
//                          gradient locally and one sum of squares per chunk;                  (reduce-scatter)
//                          its last block adds them per tensor in chunk order (no atomics: bit-reproducible) and
//                          publishes the partial ||g||^2 to every peer;
//   C  p2p_update_kernel   waits for all ranks' partials and sums them in rank order (the norms of the fully
//                          reduced gradient, identical on every rank); clip / skip / SGD / Adam on the owned range only (optimizer state is sharded: m and v
//                          of a chunk are only ever touched by its owner) and 128-bit stores of the new
//                          parameters into EVERY rank's parameter buffer; the last block publishes "params
//                          written" and waits for all peers' flags, so the kernel retires only when this
//                          rank's parameters are complete.                                        (all-gather)
//
// Flags are monotonically increasing step sequence numbers written with st.release.sys / read with
// ld.acquire.sys; all spins are bounded (a lost peer is recorded and reported by the host, never a hang or a trap).  Every rank runs the same two launches every
// step (also with an empty shard), each on its own GPU.
#pragma once
#include "elementwise.cuh"

namespace pb {
namespace p2p {

constexpr int MAX_RANKS = 8;
constexpr int MAX_TENSORS = 64;

struct SyncBlock {                        // lives in each rank's own memory, written by peers
  unsigned long long grads_ready[MAX_RANKS];
  unsigned long long norms_ready[MAX_RANKS];
  unsigned long long params_done[MAX_RANKS];
  unsigned long long pad[8];
  double norms[MAX_RANKS][MAX_TENSORS];   // partial ||g||^2 per tensor, one row per source rank
};

struct PeerTable {
  float* grads[MAX_RANKS];
  float* params[MAX_RANKS];
  SyncBlock* sync[MAX_RANKS];
  int rank, world;
  int chunk0, nchunks;                    // owned chunk range of the flat buffer
  unsigned int* lost;                     // local word: 0, or 1 + rank of a peer whose flag never arrived (host reports it)
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin on a flag a peer writes into MY memory: relaxed polls (cheap), one acquire once it flipped.  A peer that
// never arrives (crashed rank, lost link) must not hang the GPU and must not kill the context either: after ~2^25 polls
// (seconds) the waiter records 1 + the peer's rank in *lost and carries on -- every later wait of this launch and of
// the following ones sees the mark and returns at once, the step's numbers are garbage, and the host raises a clean
// error ("rank r did not arrive") when it reads the epoch's totals (read_stats).
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long seq, unsigned int* lost, int peer) {
  unsigned int polls = 0;
  while (ld_relaxed_sys(p) < seq) {
    if (++polls > 256u) __nanosleep(20);
    if ((polls & 1023u) == 0u && *reinterpret_cast<volatile unsigned int*>(lost) != 0u) return;
    if (polls > (1u << 25)) { atomicCAS(lost, 0u, (unsigned int)peer + 1u); return; }
  }
  (void)ld_acquire_sys(p);
}
// Gradients of other ranks change every step at the same addresses: never read them through the non-coherent
// path or a possibly stale L1 line -- volatile loads go to L2 / across NVLink every time.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

// ---- A: reduce-scatter + partial norms --------------------------------------------------------------------
// grid = max(1, nchunks owned), block 256.  seq_ptr[1] = number of exchanges completed since attach (never reset
// while attached, advanced by the last block of p2p_update_kernel), so flags only ever increase.
__global__ void __launch_bounds__(256) p2p_reduce_kernel(PeerTable t, const Chunk* __restrict__ chunks,
                                                         const long long* __restrict__ seq_ptr,
                                                         double* __restrict__ chunk_norm,
                                                         const int2* __restrict__ tensor_chunks, int ntensors,
                                                         unsigned int* __restrict__ blocks_done,
                                                         unsigned long long* trace) {
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)seq_ptr[1] + 1ull;   // exchanges completed since attach + 1
  if (blockIdx.x == 0 && threadIdx.x < t.world) {
    __threadfence_system();   // my backward pass (previous kernels of this stream) is visible system-wide
    st_release_sys(&t.sync[threadIdx.x]->grads_ready[t.rank], seq);
  }
  if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->grads_ready[threadIdx.x], seq, t.lost, threadIdx.x);
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(trace, 1);          // every rank's gradient is ready
  __shared__ double sh[8];
  if ((int)blockIdx.x < t.nchunks) {
    const Chunk ch = chunks[t.chunk0 + blockIdx.x];
    const int64_t n4 = ch.len >> 2;
    float4* mine = reinterpret_cast<float4*>(t.grads[t.rank] + ch.off);
    float acc = 0.f;
    // all remote loads of a pass are issued before the first add: one NVLink round trip per pass, not one per rank
    for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 2 * blockDim.x) {
      float4 v[2][MAX_RANKS];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t i = i0 + e * (int64_t)blockDim.x;
#pragma unroll
        for (int r = 0; r < MAX_RANKS; ++r)
          if (r < t.world && i < n4) v[e][r] = ld_peer(reinterpret_cast<const float4*>(t.grads[r] + ch.off) + i);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t i = i0 + e * (int64_t)blockDim.x;
        if (i < n4) {
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < MAX_RANKS; ++r)          // fixed rank order: the sum is bit-identical run to run
            if (r < t.world) { s.x += v[e][r].x; s.y += v[e][r].y; s.z += v[e][r].z; s.w += v[e][r].w; }
          mine[i] = s;
          acc = fmaf(s.x, s.x, acc); acc = fmaf(s.y, s.y, acc); acc = fmaf(s.z, s.z, acc); acc = fmaf(s.w, s.w, acc);
        }
      }
    }
    const double d = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += sh[i];
      chunk_norm[t.chunk0 + blockIdx.x] = tot;
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 2);          // owned chunks reduced
  // the last block to finish publishes this rank's per-tensor partial norms to every rank.  One fence per block,
  // after the block barrier: fences are cumulative, so it also orders the other threads' stores (a fence in every
  // thread costs a MEMBAR per warp and showed up as microseconds).
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(blocks_done, 1u) == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (last) {
    // per tensor: the owned chunks' sums of squares, added in chunk order by the whole block (independent loads, then a
    // fixed-shape reduction) -- the same bits every run -- and stored into every rank's sync block
    for (int k = 0; k < ntensors; ++k) {
      const int2 tc = tensor_chunks[k];
      const int lo = max(tc.x, t.chunk0), hi = min(tc.x + tc.y, t.chunk0 + t.nchunks);
      double a = 0.0;
      for (int i0 = lo + (int)threadIdx.x; i0 < hi; i0 += 8 * (int)blockDim.x) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + u * (int)blockDim.x < hi) ? __ldcg(&chunk_norm[i0 + u * blockDim.x]) : 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) a += v[u];
      }
      a = block_sum_d(a, sh);
      if ((int)threadIdx.x < t.world) t.sync[threadIdx.x]->norms[t.rank][k] = a;
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.world) st_release_sys(&t.sync[threadIdx.x]->norms_ready[t.rank], seq);
    if (threadIdx.x == 0) *blocks_done = 0u;
  }
  trace_end(trace);
}

// ---- C: sharded clip + update, all-gather of the new parameters ------------------------------------------
__global__ void __launch_bounds__(256) p2p_update_kernel(PeerTable t, const Chunk* __restrict__ chunks,
                                                         long long* __restrict__ seq_ptr, float* __restrict__ m,
                                                         float* __restrict__ v, UpdateParams u,
                                                         unsigned int* __restrict__ blocks_done,
                                                         unsigned long long* trace) {
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)seq_ptr[1] + 1ull;   // exchanges completed since attach + 1
  __shared__ double s_n2;
  if ((int)blockIdx.x < t.nchunks) {
    const Chunk ch = chunks[t.chunk0 + blockIdx.x];
    float scale = 1.0f;
    bool skip = false;
    if (u.rule == 0) {   // norms of the fully reduced gradient: every rank's partials, summed in rank order
      if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->norms_ready[threadIdx.x], seq, t.lost, threadIdx.x);
      __syncthreads();
      if (threadIdx.x == 0) trace_mark(trace, 1);      // every rank's partial norms are here
      if (threadIdx.x == 0) {
        double n2 = 0.0;
        for (int r = 0; r < t.world; ++r) n2 += __ldcv(&t.sync[t.rank]->norms[r][ch.tensor]);
        s_n2 = n2;
      }
      __syncthreads();
    }
    if (u.rule == 1 && u.opt == 1) {
      const double tt = (double)max((long long)u.steps_done[1], 1ll);   // Adam step count kept by the softmax kernel
      u.inv_bc1 = (float)(1.0 / (1.0 - pow(u.beta1_d, tt)));
      u.inv_bc2 = (float)(1.0 / (1.0 - pow(u.beta2_d, tt)));
    }
    if (u.rule == 0) {
      const double n2 = s_n2;
      if (!isfinite(n2)) skip = true;            // NaN/Inf gradient: this tensor keeps its value on every rank
      const double nrm = sqrt(n2);
      if (!skip && nrm > (double)u.max_norm) scale = (float)((double)u.max_norm / nrm);
    }
    if (!skip) {
      const int64_t n4 = ch.len >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(t.grads[t.rank] + ch.off);
      const float4* p4 = reinterpret_cast<const float4*>(t.params[t.rank] + ch.off);
      float4* m4 = reinterpret_cast<float4*>(m + ch.off);
      float4* v4 = reinterpret_cast<float4*>(v + ch.off);
      const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
      for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 p = p4[i], mm = make_float4(0.f, 0.f, 0.f, 0.f), vv = mm;
        if (u.opt == 1) { mm = m4[i]; vv = v4[i]; }
        adam_or_sgd4(p, mm, vv, g4[i], scale, u, omb1, omb2);
        if (u.opt == 1) { m4[i] = mm; v4[i] = vv; }
        for (int r = 0; r < t.world; ++r) *(reinterpret_cast<float4*>(t.params[r] + ch.off) + i) = p;
      }
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 2);          // owned chunks updated, parameter stores issued to every rank
  // completion: the last block of this rank publishes "my parameter writes are done" and waits for all peers
  // (one cumulative system-scope fence per block, behind the block barrier, instead of one per thread)
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(blocks_done, 1u) == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) *blocks_done = 0u;
    if (threadIdx.x < t.world) st_release_sys(&t.sync[threadIdx.x]->params_done[t.rank], seq);
    if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->params_done[threadIdx.x], seq, t.lost, threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) seq_ptr[1] = (long long)seq;   // every block of this launch has already read the old value
  }
  trace_end(trace);
}

// Optimizer state is sharded in this mode (a rank only ever touches m, v of the chunks it owns).  Before the state is
// read (checkpoints) or the exchange is left, every rank copies ITS range into a zeroed scratch buffer; an all-reduce
// (sum of disjoint shards) then gives every rank the complete state.
__global__ void owned_range_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int64_t lo, int64_t hi) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (i >= lo && i < hi) ? src[i] : 0.f;
}

}  // namespace p2p
}  // namespace pb
