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
//                          gradient locally and accumulates per-tensor ||g||^2 partials;        (reduce-scatter)
//                          its last block publishes the partials to every peer;
//   C  p2p_update_kernel   waits for all ranks' partials and sums them in rank order (the norms of the fully
//                          reduced gradient, identical on every rank); clip / skip / SGD / Adam on the owned range only (optimizer state is sharded: m and v
//                          of a chunk are only ever touched by its owner) and 128-bit stores of the new
//                          parameters into EVERY rank's parameter buffer; the last block publishes "params
//                          written" and waits for all peers' flags, so the kernel retires only when this
//                          rank's parameters are complete.                                        (all-gather)
//
// Flags are monotonically increasing step sequence numbers written with st.release.sys / read with
// ld.acquire.sys; all spins are bounded (trap, never hang).  Every rank runs the same two launches every
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
// bounded spin on a flag a peer writes into MY memory: relaxed polls (cheap), one acquire once it flipped;
// ~2^26 polls (seconds) and a lost peer traps instead of hanging the GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* p, unsigned long long seq) {
  unsigned int polls = 0;
  while (ld_relaxed_sys(p) < seq) {
    if (++polls > 256u) __nanosleep(20);
    if (polls > (1u << 26)) __trap();
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
                                                         double* __restrict__ norm_part, int ntensors,
                                                         unsigned int* __restrict__ blocks_done,
                                                         unsigned long long* trace) {
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)seq_ptr[1] + 1ull;   // exchanges completed since attach + 1
  if (blockIdx.x == 0 && threadIdx.x < t.world) {
    __threadfence_system();   // my backward pass (previous kernels of this stream) is visible system-wide
    st_release_sys(&t.sync[threadIdx.x]->grads_ready[t.rank], seq);
  }
  if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->grads_ready[threadIdx.x], seq);
  __syncthreads();
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
      atomicAdd(&norm_part[ch.tensor], tot);
    }
  }
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
    for (int i = threadIdx.x; i < t.world * ntensors; i += blockDim.x) {
      const int r = i / ntensors, k = i - r * ntensors;
      t.sync[r]->norms[t.rank][k] = __ldcg(&norm_part[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.world) st_release_sys(&t.sync[threadIdx.x]->norms_ready[t.rank], seq);
    for (int k = threadIdx.x; k < ntensors; k += blockDim.x) norm_part[k] = 0.0;   // ready for the next step
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
      if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->norms_ready[threadIdx.x], seq);
      __syncthreads();
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
        const float4 p = p4[i], g = g4[i];
        float pe[4] = {p.x, p.y, p.z, p.w};
        const float ge[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
        if (u.opt == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pe[k] -= u.lr * ge[k];
        } else {
          const float4 mm = m4[i], vv = v4[i];
          float me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            me[k] = me[k] * u.beta1 + omb1 * ge[k];
            ve[k] = ve[k] * u.beta2 + omb2 * (ge[k] * ge[k]);
            const float mh = me[k] * u.inv_bc1;
            float vh = ve[k] * u.inv_bc2;
            if (u.rule == 0) vh = fmaxf(vh, u.eps2);
            pe[k] -= u.lr * mh / (sqrtf(vh) + u.eps);
          }
          m4[i] = make_float4(me[0], me[1], me[2], me[3]);
          v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        }
        const float4 np = make_float4(pe[0], pe[1], pe[2], pe[3]);
        for (int r = 0; r < t.world; ++r) *(reinterpret_cast<float4*>(t.params[r] + ch.off) + i) = np;
      }
    }
  }
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
    if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->params_done[threadIdx.x], seq);
    __syncthreads();
    if (threadIdx.x == 0) seq_ptr[1] = (long long)seq;   // every block of this launch has already read the old value
  }
  trace_end(trace);
}

// ---- A + C in one launch (opt-in, PYCNN_B200_P2P_FUSED=1) ---------------------------------------------------
// Same protocol, one kernel: a block keeps its reduced chunk in registers across the norm exchange, so the reduced
// gradient is not re-read and one launch boundary disappears.  Every block waits (step 4) for flags that are only
// posted once ALL blocks of every rank have finished step 2, so the whole grid must be co-resident: the host
// launches this kernel only when nchunks fits the device (chunks are 2048 floats = two float4 per thread).
__global__ void __launch_bounds__(256, 2) p2p_fused_kernel(PeerTable t, const Chunk* __restrict__ chunks,
                                                           long long* __restrict__ seq_ptr,
                                                           double* __restrict__ norm_part, int ntensors,
                                                           float* __restrict__ m, float* __restrict__ v, UpdateParams u,
                                                           unsigned int* __restrict__ blocks_done,
                                                           unsigned long long* trace) {
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)seq_ptr[1] + 1ull;
  // 1. gradients of every rank are complete
  if (blockIdx.x == 0 && threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(&t.sync[threadIdx.x]->grads_ready[t.rank], seq);
  }
  if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->grads_ready[threadIdx.x], seq);
  __syncthreads();
  // 2. reduce-scatter of the owned chunk into registers (fixed rank order), partial norm
  __shared__ double sh[8];
  __shared__ double s_n2;
  __shared__ bool last;
  const bool owns = (int)blockIdx.x < t.nchunks;
  Chunk ch{};
  float4 red[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  int64_t n4 = 0;
  if (owns) {
    ch = chunks[t.chunk0 + blockIdx.x];
    n4 = ch.len >> 2;
    float4 part[2][MAX_RANKS];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t i = threadIdx.x + e * 256;
#pragma unroll
      for (int r = 0; r < MAX_RANKS; ++r)
        if (r < t.world && i < n4) part[e][r] = ld_peer(reinterpret_cast<const float4*>(t.grads[r] + ch.off) + i);
    }
    float acc = 0.f;
    float4* mine = reinterpret_cast<float4*>(t.grads[t.rank] + ch.off);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int64_t i = threadIdx.x + e * 256;
      if (i < n4) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < MAX_RANKS; ++r)
          if (r < t.world) { sum.x += part[e][r].x; sum.y += part[e][r].y; sum.z += part[e][r].z; sum.w += part[e][r].w; }
        red[e] = sum;
        mine[i] = sum;     // get_grad() after a step still returns the reduced gradient
        acc = fmaf(sum.x, sum.x, acc); acc = fmaf(sum.y, sum.y, acc); acc = fmaf(sum.z, sum.z, acc); acc = fmaf(sum.w, sum.w, acc);
      }
    }
    const double d = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += sh[i];
      atomicAdd(&norm_part[ch.tensor], tot);
    }
  }
  // 3. the last block publishes this rank's partial norms to every rank
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(blocks_done, 1u) == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (last) {
    for (int i = threadIdx.x; i < t.world * ntensors; i += blockDim.x) {
      const int r = i / ntensors, k = i - r * ntensors;
      t.sync[r]->norms[t.rank][k] = __ldcg(&norm_part[k]);
    }
    __syncthreads();
    if (threadIdx.x == 0) __threadfence_system();
    __syncthreads();
    if (threadIdx.x < t.world) st_release_sys(&t.sync[threadIdx.x]->norms_ready[t.rank], seq);
    for (int k = threadIdx.x; k < ntensors; k += blockDim.x) norm_part[k] = 0.0;
    if (threadIdx.x == 0) *blocks_done = 0u;
  }
  // 4.-5. norms of the fully reduced gradient, sharded clip + update from registers, all-gather of the parameters
  if (owns) {
    float scale = 1.0f;
    bool skip = false;
    if (u.rule == 0) {
      if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->norms_ready[threadIdx.x], seq);
      __syncthreads();
      if (threadIdx.x == 0) {
        double n2 = 0.0;
        for (int r = 0; r < t.world; ++r) n2 += __ldcv(&t.sync[t.rank]->norms[r][ch.tensor]);
        s_n2 = n2;
      }
      __syncthreads();
      const double n2 = s_n2;
      if (!isfinite(n2)) skip = true;
      const double nrm = sqrt(n2);
      if (!skip && nrm > (double)u.max_norm) scale = (float)((double)u.max_norm / nrm);
    }
    if (u.rule == 1 && u.opt == 1) {
      const double tt = (double)max((long long)u.steps_done[1], 1ll);
      u.inv_bc1 = (float)(1.0 / (1.0 - pow(u.beta1_d, tt)));
      u.inv_bc2 = (float)(1.0 / (1.0 - pow(u.beta2_d, tt)));
    }
    if (!skip) {
      const float4* p4 = reinterpret_cast<const float4*>(t.params[t.rank] + ch.off);
      float4* m4 = reinterpret_cast<float4*>(m + ch.off);
      float4* v4 = reinterpret_cast<float4*>(v + ch.off);
      const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t i = threadIdx.x + e * 256;
        if (i >= n4) continue;
        const float4 p = p4[i];
        float pe[4] = {p.x, p.y, p.z, p.w};
        const float ge[4] = {red[e].x * scale, red[e].y * scale, red[e].z * scale, red[e].w * scale};
        if (u.opt == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) pe[k] -= u.lr * ge[k];
        } else {
          const float4 mm = m4[i], vv = v4[i];
          float me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            me[k] = me[k] * u.beta1 + omb1 * ge[k];
            ve[k] = ve[k] * u.beta2 + omb2 * (ge[k] * ge[k]);
            const float mh = me[k] * u.inv_bc1;
            float vh = ve[k] * u.inv_bc2;
            if (u.rule == 0) vh = fmaxf(vh, u.eps2);
            pe[k] -= u.lr * mh / (sqrtf(vh) + u.eps);
          }
          m4[i] = make_float4(me[0], me[1], me[2], me[3]);
          v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
        }
        const float4 np = make_float4(pe[0], pe[1], pe[2], pe[3]);
        for (int r = 0; r < t.world; ++r) *(reinterpret_cast<float4*>(t.params[r] + ch.off) + i) = np;
      }
    }
  }
  // 6. completion barrier across ranks
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(blocks_done + 1, 1u) == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (last) {
    if (threadIdx.x == 0) blocks_done[1] = 0u;
    if (threadIdx.x < t.world) st_release_sys(&t.sync[threadIdx.x]->params_done[t.rank], seq);
    if (threadIdx.x < t.world) wait_flag(&t.sync[t.rank]->params_done[threadIdx.x], seq);
    __syncthreads();
    if (threadIdx.x == 0) seq_ptr[1] = (long long)seq;
  }
  trace_end(trace);
}

}  // namespace p2p
}  // namespace pb
