// elementwise.cuh -- HBM-bound kernels of the MLP path: row gather, fused softmax+CE+argmax+dZ,
// bias column sums, per-tensor gradient norms and the clipped SGD/Adam update, converts.
#pragma once
#include "common.cuh"

namespace pb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses that do not pollute L1 (guide: guideline 13)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// K8: batch_x = data[batch_idx], batch_y = labels[batch_idx]   (pycnn/pycnn.py:629-630)
// rows are Dp floats (Dp % 4 == 0, 16-byte aligned): one float4 per thread, grid-strided.
// ---------------------------------------------------------------------------------------------
// The batch's indices are idx_base[*cursor + extra ...]: the cursor lives on the device so that one captured
// CUDA graph can be replayed for every batch of an epoch (advance_cursor_kernel moves it).
template <class I>
__device__ __forceinline__ void gather_rows_body(const float* __restrict__ feat, const int32_t* __restrict__ labels,
                                                 const int64_t* __restrict__ idx, I total, I vec_per_row, int64_t Dp,
                                                 float* __restrict__ x, int32_t* __restrict__ blabels) {
  constexpr int U = 4;   // independent 128-bit loads in flight per thread: the kernel also runs with few CTAs per SM
  const I stride = (I)gridDim.x * blockDim.x;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride * U) {
    float4 v[U];
    I r[U], c[U];
    int64_t src[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const I iu = i + (I)u * stride;
      if (iu < total) {
        r[u] = iu / vec_per_row; c[u] = iu - r[u] * vec_per_row;
        src[u] = idx[r[u]];
        v[u] = ld_stream(reinterpret_cast<const float4*>(feat + src[u] * Dp) + c[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const I iu = i + (I)u * stride;
      if (iu < total) {
        st_stream(reinterpret_cast<float4*>(x + (int64_t)r[u] * Dp) + c[u], v[u]);
        if (c[u] == 0) blabels[r[u]] = labels[src[u]];
      }
    }
  }
}

// Rows of at most 5 x 256 float4 (D <= 5120): a CTA moves two whole rows per iteration -- one index load and one
// base pointer per row, up to 10 independent 128-bit loads in flight per thread with little address state, so the
// kernel keeps HBM busy even at the 2 CTAs per SM it gets on the prefetch stream.
__device__ __forceinline__ void gather_rows_short(const float* __restrict__ feat, const int32_t* __restrict__ labels,
                                                  const int64_t* __restrict__ idx, int bs, int vec_per_row, int64_t Dp,
                                                  float* __restrict__ x, int32_t* __restrict__ blabels) {
  constexpr int NV = 5, RPI = 2;
  for (int row = blockIdx.x * RPI; row < bs; row += gridDim.x * RPI) {
    float4 v[RPI][NV];
    int64_t src[RPI];
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      src[r] = (row + r < bs) ? idx[row + r] : -1;
      const float4* s4 = reinterpret_cast<const float4*>(feat + (src[r] < 0 ? 0 : src[r]) * Dp);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = threadIdx.x + j * 256;
        if (src[r] >= 0 && c < vec_per_row) v[r][j] = ld_stream(s4 + c);
      }
    }
#pragma unroll
    for (int r = 0; r < RPI; ++r) {
      if (src[r] < 0) continue;
      float4* d4 = reinterpret_cast<float4*>(x + (int64_t)(row + r) * Dp);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int c = threadIdx.x + j * 256;
        if (c < vec_per_row) st_stream(d4 + c, v[r][j]);
      }
      if (threadIdx.x == 0) blabels[row + r] = labels[src[r]];
    }
  }
}

__global__ void __launch_bounds__(256, 4) gather_rows_kernel(const float* __restrict__ feat,
                                                          const int32_t* __restrict__ labels,
                                                          const int64_t* __restrict__ idx_base,
                                                          const int64_t* __restrict__ cursor, int64_t extra, int bs,
                                                          int64_t Dp, float* __restrict__ x,
                                                          int32_t* __restrict__ blabels,
                                                          unsigned long long* trace) {
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(trace);
  const int64_t* __restrict__ idx = idx_base + (cursor ? cursor[0] : 0) + extra;
  const int64_t vec_per_row = Dp >> 2;
  const int64_t total = (int64_t)bs * vec_per_row;
  if (vec_per_row <= 5 * 256 && vec_per_row > 2 * 256)
    gather_rows_short(feat, labels, idx, bs, (int)vec_per_row, Dp, x, blabels);
  else if (total < (int64_t)0x7fffffff - (int64_t)4 * gridDim.x * blockDim.x)
    gather_rows_body<uint32_t>(feat, labels, idx, (uint32_t)total, (uint32_t)vec_per_row, Dp, x, blabels);
  else
    gather_rows_body<int64_t>(feat, labels, idx, total, vec_per_row, Dp, x, blabels);
  trace_end(trace);
}

// Zero-copy batch assembly: rows idx[i] of a pinned host array (device-visible through UVA) are pulled over
// PCIe with 128-bit loads straight into the device staging buffer -- no host-side gather, no copy-engine
// round trip per row.  row_bytes % 16 == 0 is guaranteed by the caller (else the byte tail loop runs).
__global__ void __launch_bounds__(256) gather_host_rows_kernel(const uint8_t* __restrict__ host_rows,
                                                               const int32_t* __restrict__ host_labels,
                                                               const int64_t* __restrict__ idx, int bs,
                                                               int64_t row_bytes, uint8_t* __restrict__ dst,
                                                               int32_t* __restrict__ dst_labels) {
  const int64_t vec = row_bytes >> 4, tail = row_bytes & 15;
  const int64_t total = (int64_t)bs * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / vec, cidx = i - r * vec;
    const int64_t src = idx[r];
    const uint4 v = *(reinterpret_cast<const uint4*>(host_rows + src * row_bytes) + cidx);
    *(reinterpret_cast<uint4*>(dst + r * row_bytes) + cidx) = v;
    if (cidx == 0) dst_labels[r] = host_labels[src];
  }
  if (tail) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)bs * tail;
         i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / tail, b = vec * 16 + (i - r * tail);
      dst[r * row_bytes + b] = host_rows[idx[r] * row_bytes + b];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K4+K5+K6(a): row softmax (max-subtracted, forward_pass.pyx:40-41), CE loss
// -log(p[y]+1e-9) (optimized.cpp:34-46), argmax==label count (pycnn.py:637-639) and
// dZ = p - onehot (backward_pass.pyx:65) in ONE pass; one warp per row, shuffles only.
// logits/probs/dz are [bs, ld]; padding columns of dz are written as 0.
// stats[0] += sum of row losses (double); stats[1] += #correct (as a double, exact below 2^53).
// Any of probs/dz/pred/conf may be null.  labels may be null (inference: no loss/dz).
// ---------------------------------------------------------------------------------------------
struct StepBookkeeping {   // per-step housekeeping folded into the softmax kernel / epilogue (block 0) to save graph nodes
  int64_t* cursor;         // cursor[0] += advance (next batch of the permutation); cursor[1] += 1 (step count t)
  int64_t advance;
};
__device__ __forceinline__ void step_bookkeeping(const StepBookkeeping& bk) {   // one thread of the launch
  // advance == 0: the row-gather prefetch branch owns cursor[0] (advance_cursor_kernel on pf_stream); a
  // read-modify-write of +0 here could swallow its advance
  if (bk.cursor) { if (bk.advance) bk.cursor[0] += bk.advance; bk.cursor[1] += 1; }
}

// Running (loss, #correct) totals are integers so that their atomic accumulation is order-independent and a
// training run reproduces bit for bit: stats[0] = sum of row losses in 2^-32 fixed point (two's complement),
// stats[1] = #correct, stats[2] = number of non-finite contributions (the host then reports NaN, as the
// reference's float sum would), stats[3] unused.  |loss sum| < 2^31 covers 10^8 rows at the -log(1e-9) worst case.
constexpr int kStatsWords = 4;
constexpr double kLossFixedScale = 4294967296.0;
__device__ __forceinline__ void stats_add(unsigned long long* stats, double loss, unsigned long long correct) {
  if (!isfinite(loss)) atomicAdd(&stats[2], 1ull);
  else atomicAdd(&stats[0], (unsigned long long)__double2ll_rn(loss * kLossFixedScale));
  if (correct) atomicAdd(&stats[1], correct);
}

// gather-free steps: only the labels of the batch are gathered (bs x 4 bytes); the feature rows are fetched by the layer-1
// GEMMs themselves
__global__ void gather_labels_kernel(const int32_t* __restrict__ labels, const int64_t* __restrict__ idx_base,
                                     const int64_t* __restrict__ cursor, int bs, int32_t* __restrict__ blabels) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t* idx = idx_base + (cursor ? cursor[0] : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < bs; i += gridDim.x * blockDim.x) blabels[i] = labels[idx[i]];
}

__global__ void advance_cursor_kernel(int64_t* cursor, int64_t advance) {   // tail of the prefetch branch
  if (threadIdx.x == 0) cursor[0] += advance;
}

__global__ void step_bookkeeping_kernel(StepBookkeeping bk) {   // ranks whose shard of a batch is empty
  pdl_wait();
  if (threadIdx.x == 0) step_bookkeeping(bk);
}

// One warp per row, rows grid-strided (a warp visits its rows in increasing order).  db_part (training): the bias
// gradient of the output layer as ORDERED partial sums -- block b writes db_part[b * ld + j] = sum over its rows of
// dZ[., j] (each warp accumulates its own rows in a private smem row, the 8 warp rows are then added in warp order),
// and grad_finalize_kernel adds the gridDim.x partials in block order: no atomics, bit-reproducible.
__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits, int bs, int C,
                                                         int64_t ld, const int32_t* __restrict__ labels,
                                                         float* __restrict__ probs, float* __restrict__ dz,
                                                         int32_t* __restrict__ pred, float* __restrict__ conf,
                                                         unsigned long long* __restrict__ stats,
                                                         float* __restrict__ db_part, StepBookkeeping bk) {
  extern __shared__ float s_db[];   // [8][C] when db_part != nullptr
  pdl_wait();
  pdl_launch_dependents();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0) step_bookkeeping(bk);
  if (db_part)
    for (int j = lane; j < C; j += 32) s_db[w * C + j] = 0.f;
  double my_loss = 0.0;
  unsigned long long my_correct = 0ull;
  for (int row = blockIdx.x * warps_per_block + w; row < bs; row += gridDim.x * warps_per_block) {
    const float* z = logits + (int64_t)row * ld;
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = lane; j < C; j += 32) {
      const float v = z[j];
      if (v > mx) { mx = v; arg = j; }   // first maximum within the lane's strided subsequence
    }
    // warp arg-max with lowest-index tie-break (np.argmax semantics)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oarg = __shfl_xor_sync(0xffffffffu, arg, o);
      if (omx > mx || (omx == mx && oarg < arg)) { mx = omx; arg = oarg; }
    }
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += expf(z[j] - mx);
    s = warp_sum(s);
    const float inv = 1.0f / s;
    const int y = labels ? labels[row] : -1;
    for (int j = lane; j < (int)ld; j += 32) {
      float p = 0.f;
      if (j < C) p = expf(z[j] - mx) * inv;
      if (probs) probs[(int64_t)row * ld + j] = p;
      const float d = (j < C) ? (p - (j == y ? 1.0f : 0.0f)) : 0.0f;
      if (dz) dz[(int64_t)row * ld + j] = d;
      if (db_part && j < C) s_db[w * C + j] += d;          // db_o = sum_rows dZ (backward_pass.pyx:72)
      if (j == y) my_loss += -log((double)p + 1e-9);
    }
    if (lane == 0) {
      if (pred) pred[row] = arg;
      if (conf) conf[row] = inv;  // p[argmax] = exp(0)/s
      if (y >= 0 && arg == y) my_correct += 1ull;
    }
  }
  if (db_part) {
    __syncthreads();
    for (int j = threadIdx.x; j < (int)ld; j += blockDim.x) {
      float t = 0.f;
      if (j < C)
        for (int i = 0; i < warps_per_block; ++i) t += s_db[i * C + j];
      db_part[(int64_t)blockIdx.x * ld + j] = t;
    }
  }
  if (stats && labels) {
    // block-level reduction (fixed order) -> one integer atomic pair per block
    __shared__ double s_loss[8];
    __shared__ unsigned long long s_corr[8];
    my_loss = warp_sum_d(my_loss);
    if (lane == 0) { s_loss[w] = my_loss; s_corr[w] = my_correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double l = 0.0;
      unsigned long long c = 0ull;
      for (int i = 0; i < warps_per_block; ++i) { l += s_loss[i]; c += s_corr[i]; }
      stats_add(stats, l, c);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K6(c): bias gradients db[n] = sum_rows dZ[m, n]  (backward_pass.pyx:72,93) for the FP32-SIMT path, as ORDERED
// partial sums: grid (ceil(N/32), R); block (32, 8): thread (x, y) sums rows y, y+8, ... of row chunk blockIdx.y
// for column x in FP64, the 8 partials are combined through smem in y order and the block writes
// part[blockIdx.y * ld_part + n]; grad_finalize_kernel adds the R partials in chunk order (no atomics).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, int bs, int N, int64_t ld,
                                                     float* __restrict__ part, int64_t ld_part, int rows_per_block) {
  __shared__ double red[8][33];
  pdl_wait();
  pdl_launch_dependents();
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(bs, r0 + rows_per_block);
  double acc = 0.0;
  if (n < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) acc += (double)dz[(int64_t)r * ld + n];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    part[(int64_t)blockIdx.y * ld_part + n] = (float)s;
  }
}

// ---------------------------------------------------------------------------------------------
// Gradient accumulation is ORDERED, never atomic: a producer that cannot finish a gradient tensor alone (split-K
// dW GEMMs: one partial per k-slice; bias gradients: one partial per row tile) writes its partial sums side by side
// into an arena, and grad_finalize_kernel adds them in index order into the flat gradient buffer.  Same inputs ->
// bit-identical gradients, norms and parameters, run after run (backward_pass.pyx:71-72,92-93 are plain sums).
// ---------------------------------------------------------------------------------------------
struct GradParts {               // by-value kernel argument
  static constexpr int MAX_T = 64;
  const float* base;             // partial-sum arena
  int64_t off[MAX_T];            // offset of a tensor's partial 0 in the arena
  int64_t stride[MAX_T];         // distance between consecutive partials
  int nparts[MAX_T];             // 0: the producer wrote the flat gradient buffer directly
};

// fixed-shape block reduction of one double per thread (256 threads): identical result in every thread
__device__ __forceinline__ double block_sum_d(double v, double* s8) {
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) s8[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += s8[i];
  __syncthreads();
  return t;
}

// One block per chunk of the flat buffer: grads[chunk] = sum_p part[p][chunk] (when the tensor has partials) and,
// if asked, chunk_norm[chunk] = sum of squares (gradient_decent.pyx:33, np.linalg.norm per tensor -- the per-tensor
// total is formed by the update kernel from the chunk values in chunk order).
// Tensors that can collect many partials (biases: one per row tile; small weight matrices: up to 32 k-slices) are cut
// into short chunks (kSmallChunk floats, set_network), and their blocks split the partials over the 8 warps: warp w
// adds partials [w*G, (w+1)*G) in order, the 8 group sums meet in smem in warp order.  The summation tree depends only
// on (np, chunk shape): bit-reproducible.
constexpr int kSmallChunk = 512;   // floats
__global__ void __launch_bounds__(256) grad_finalize_kernel(float* __restrict__ grads, const Chunk* __restrict__ chunks,
                                                            const GradParts gp, double* __restrict__ chunk_norm,
                                                            int want_norm, unsigned long long* trace) {
  __shared__ double s8[8];
  __shared__ float4 group_sum[8][kSmallChunk / 4];
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(trace);
  const Chunk ch = chunks[blockIdx.x];
  const int np = gp.nparts[ch.tensor];
  float4* g4 = reinterpret_cast<float4*>(grads + ch.off);
  const int n4 = (int)(ch.len >> 2);  // chunk offsets/lengths are multiples of 4
  float acc = 0.f;
  if (np > 8 && n4 <= kSmallChunk / 4) {
    const float4* p0 = reinterpret_cast<const float4*>(gp.base + gp.off[ch.tensor] + ch.rel);
    const int64_t stride4 = gp.stride[ch.tensor] >> 2;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = (np + 7) >> 3;
    const int pb = w * G, pe = min(np, pb + G);
    float4 s[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = pb; p < pe; p += 2) {          // 8 independent loads in flight per lane
      float4 v[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (p + u < pe && lane + 32 * j < n4) v[u][j] = __ldcg(p0 + (int64_t)(p + u) * stride4 + lane + 32 * j);
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (p + u < pe && lane + 32 * j < n4) { s[j].x += v[u][j].x; s[j].y += v[u][j].y; s[j].z += v[u][j].z; s[j].w += v[u][j].w; }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lane + 32 * j < n4) group_sum[w][lane + 32 * j] = s[j];
    __syncthreads();
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 t = group_sum[0][i];
#pragma unroll
      for (int k = 1; k < 8; ++k) { const float4 o = group_sum[k][i]; t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w; }
      g4[i] = t;
      acc = fmaf(t.x, t.x, acc); acc = fmaf(t.y, t.y, acc); acc = fmaf(t.z, t.z, acc); acc = fmaf(t.w, t.w, acc);
    }
  } else if (np > 0) {
    const float* p0 = gp.base + gp.off[ch.tensor] + ch.rel;
    const int64_t stride = gp.stride[ch.tensor];
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < np; p += 8) {      // 8 independent loads in flight, added in index order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (p + u < np) v[u] = __ldcg(reinterpret_cast<const float4*>(p0 + (int64_t)(p + u) * stride) + i);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (p + u < np) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
      }
      g4[i] = s;
      acc = fmaf(s.x, s.x, acc); acc = fmaf(s.y, s.y, acc); acc = fmaf(s.z, s.z, acc); acc = fmaf(s.w, s.w, acc);
    }
  } else if (want_norm) {
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = g4[i];
      acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
  }
  if (want_norm) {
    const double t = block_sum_d((double)acc, s8);
    if (threadIdx.x == 0) chunk_norm[blockIdx.x] = t;
  }
  trace_end(trace);
}

// norms of a gradient that is already complete in the flat buffer (after an NCCL all-reduce / external exchange)
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ grads,
                                                          const Chunk* __restrict__ chunks,
                                                          double* __restrict__ chunk_norm, unsigned long long* trace) {
  __shared__ double s8[8];
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(trace);
  const Chunk ch = chunks[blockIdx.x];
  const float4* g4 = reinterpret_cast<const float4*>(grads + ch.off);
  const int64_t n4 = ch.len >> 2;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
  }
  const double t = block_sum_d((double)acc, s8);
  if (threadIdx.x == 0) chunk_norm[blockIdx.x] = t;
  trace_end(trace);
}

struct UpdateParams {
  const double* tensor_norm;   // per-tensor ||g||^2 formed by tensor_norm_kernel, or nullptr: every block adds its tensor's chunk
                               // values itself (cheaper than a launch for small networks, quadratic for large ones)
  int opt;       // 0 sgd, 1 adam
  int rule;      // 0 cpu (clip + skip), 1 cupy
  float lr, beta1, beta2, eps, inv_bc1, inv_bc2, eps2, max_norm;
  double beta1_d, beta2_d;
  const int64_t* steps_done;  // device cursor; [1] = Adam step count t of the current step (CuPy rule)
};

// ||g||^2 of the chunk's tensor: its chunk values added in chunk order by a fixed-shape block reduction, so every
// block of a tensor (and every run) arrives at the same bits
__device__ __forceinline__ double tensor_norm2(const double* __restrict__ chunk_norm, const Chunk& ch, double* s8) {
  double a = 0.0;
  for (int i0 = threadIdx.x; i0 < ch.n_chunks; i0 += 4 * blockDim.x) {      // independent L2 loads, added in chunk order
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (i0 + u * (int)blockDim.x < ch.n_chunks) ? __ldcg(chunk_norm + ch.first_chunk + i0 + u * blockDim.x) : 0.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) a += v[u];
  }
  return block_sum_d(a, s8);
}

// large networks: one block per tensor adds the tensor's chunk values in chunk order (fixed-shape reduction)
__global__ void __launch_bounds__(256) tensor_norm_kernel(const double* __restrict__ chunk_norm, const int2* __restrict__ tensor_chunks,
                                                          double* __restrict__ tensor_norm) {
  __shared__ double s8[8];
  pdl_wait();
  pdl_launch_dependents();
  const int2 tc = tensor_chunks[blockIdx.x];
  double a = 0.0;
  for (int i0 = threadIdx.x; i0 < tc.y; i0 += 8 * blockDim.x) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (i0 + u * (int)blockDim.x < tc.y) ? __ldcg(chunk_norm + tc.x + i0 + u * blockDim.x) : 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) a += v[u];
  }
  const double t = block_sum_d(a, s8);
  if (threadIdx.x == 0) tensor_norm[blockIdx.x] = t;
}

__device__ __forceinline__ void adam_or_sgd4(float4& p, float4& mm, float4& vv, const float4& g, float scale,
                                             const UpdateParams& u, float omb1, float omb2) {
  float pe[4] = {p.x, p.y, p.z, p.w}, me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
  const float ge[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
  if (u.opt == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) pe[k] -= u.lr * ge[k];
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      me[k] = me[k] * u.beta1 + omb1 * ge[k];
      ve[k] = ve[k] * u.beta2 + omb2 * (ge[k] * ge[k]);
      const float mh = me[k] * u.inv_bc1;
      float vh = ve[k] * u.inv_bc2;
      if (u.rule == 0) vh = fmaxf(vh, u.eps2);
      pe[k] -= u.lr * mh / (sqrtf(vh) + u.eps);
    }
    mm = make_float4(me[0], me[1], me[2], me[3]);
    vv = make_float4(ve[0], ve[1], ve[2], ve[3]);
  }
  p = make_float4(pe[0], pe[1], pe[2], pe[3]);
}

// clipped SGD / Adam over the flat buffers; each block handles one chunk and forms its tensor's norm.
// CPU rule (gradient_decent.pyx): g *= 5/||g|| when ||g|| > 5 (:33-35); non-finite -> skip tensor
// (:37-39); Adam m,v in place (:91-95), v_hat floored at eps^2 (:98), w -= lr*m_hat/(sqrt(v_hat)+eps).
__global__ void __launch_bounds__(256) clip_update_kernel(float* __restrict__ params,
                                                          const float* __restrict__ grads,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const Chunk* __restrict__ chunks,
                                                          const double* __restrict__ chunk_norm, UpdateParams u,
                                                          unsigned long long* trace) {
  __shared__ double s8[8];
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(trace);
  const Chunk ch = chunks[blockIdx.x];
  float scale = 1.0f;
  if (u.rule == 1 && u.opt == 1) {   // pycnn/pycnn.py:534,544-545: t advances, 1 - beta^t
    const double t = (double)max((long long)u.steps_done[1], 1ll);   // incremented by this step's softmax kernel
    u.inv_bc1 = (float)(1.0 / (1.0 - pow(u.beta1_d, t)));
    u.inv_bc2 = (float)(1.0 / (1.0 - pow(u.beta2_d, t)));
  }
  // a chunk is at most 2048 floats = two 128-bit elements per thread: all of the thread's operands are requested
  // BEFORE the tensor norm is formed, so the HBM round trip of p, g, m, v overlaps the norm's L2 round trip + reduction
  const int n4 = (int)(ch.len >> 2);
  float4* p4 = reinterpret_cast<float4*>(params + ch.off);
  const float4* g4 = reinterpret_cast<const float4*>(grads + ch.off);
  float4* m4 = reinterpret_cast<float4*>(m + ch.off);
  float4* v4 = reinterpret_cast<float4*>(v + ch.off);
  float4 p[2], g[2], mm[2], vv[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = threadIdx.x + e * 256;
    mm[e] = vv[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      p[e] = p4[i]; g[e] = __ldcg(g4 + i);
      if (u.opt == 1) { mm[e] = m4[i]; vv[e] = v4[i]; }
    }
  }
  if (u.rule == 0) {
    const double n2 = u.tensor_norm ? __ldcg(u.tensor_norm + ch.tensor) : tensor_norm2(chunk_norm, ch, s8);
    if (!isfinite(n2)) return;  // NaN/Inf gradient: skip this tensor's update
    const double nrm = sqrt(n2);
    if (nrm > (double)u.max_norm) scale = (float)((double)u.max_norm / nrm);
  }
  const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = threadIdx.x + e * 256;
    if (i < n4) {
      adam_or_sgd4(p[e], mm[e], vv[e], g[e], scale, u, omb1, omb2);
      p4[i] = p[e];
      if (u.opt == 1) { m4[i] = mm[e]; v4[i] = vv[e]; }
    }
  }
  trace_end(trace);
}

// ---------------------------------------------------------------------------------------------
// converts between host-facing float64 matrices [rows, cols] (dense) and device float32 [rows, ld]
// ---------------------------------------------------------------------------------------------
__global__ void f64_to_f32_padded_kernel(const double* __restrict__ src, int64_t rows, int64_t cols,
                                         float* __restrict__ dst, int64_t ld) {
  const int64_t total = rows * ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld, c = i - r * ld;
    dst[i] = (c < cols) ? (float)src[r * cols + c] : 0.0f;
  }
}
__global__ void f32_padded_to_f64_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                         double* __restrict__ dst) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    dst[i] = (double)src[r * ld + c];
  }
}
__global__ void f32_padded_to_f32_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld,
                                         float* __restrict__ dst) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    dst[i] = src[r * ld + c];
  }
}

// split-K without atomics (k-slices that cannot meet in distributed shared memory): out = epilogue(sum of the S raw
// partial tiles in slice order).  epi: 0 none, 1 +bias, 2 relu(+bias), 3 mask by (act > 0).  128-bit accesses: ldc and
// part_stride are multiples of 4 and the partial rows carry exact zeros in their padding columns.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int S, int64_t part_stride,
                                                            float* __restrict__ c, int M, int N, int64_t ldc, int epi,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ mask, int64_t ldmask) {
  pdl_wait();
  pdl_launch_dependents();
  const int n4 = (N + 3) >> 2;
  const int64_t total = (int64_t)M * n4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n4;
    const int n = (int)(i - r * n4) * 4;
    const float* p0 = ws + r * ldc + n;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < S; z += 4) {          // 4 independent loads in flight, added in slice order
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (z + u < S) t[u] = __ldcg(reinterpret_cast<const float4*>(p0 + (int64_t)(z + u) * part_stride));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (z + u < S) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
    }
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (n + e >= N) { o[e] = 0.f; continue; }
      if (epi == 1 || epi == 2) o[e] += __ldg(bias + n + e);
      if (epi == 2) o[e] = (o[e] < 0.f) ? 0.f : o[e];
      if (epi == 3) o[e] = (__ldg(mask + r * ldmask + n + e) <= 0.f) ? 0.f : o[e];
    }
    *reinterpret_cast<float4*>(c + r * ldc + n) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// standalone drop-ins for optimized.cpp on float64 data (tiny; used by _softmax_batch/_cross_entropy_batch)
__global__ void softmax_f64_kernel(const double* __restrict__ x, double* __restrict__ out, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const double* z = x + (int64_t)row * cols;
  double mx = -INFINITY;
  for (int j = lane; j < cols; j += 32) mx = fmax(mx, z[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  double s = 0.0;
  for (int j = lane; j < cols; j += 32) s += exp(z[j] - mx);
  s = warp_sum_d(s);
  const double inv = 1.0 / s;
  for (int j = lane; j < cols; j += 32) out[(int64_t)row * cols + j] = exp(z[j] - mx) * inv;
}
__global__ void cross_entropy_f64_kernel(const double* __restrict__ p, const double* __restrict__ y,
                                         double* __restrict__ out, int rows, int cols) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  double s = 0.0;
  for (int j = lane; j < cols; j += 32)
    s += y[(int64_t)row * cols + j] * log(p[(int64_t)row * cols + j] + 1e-9);
  s = warp_sum_d(s);
  if (lane == 0) out[row] = -s;
}
__global__ void max_pool_f64_kernel(const double* __restrict__ fm, int h, int w, double* __restrict__ out) {
  const int nh = h / 2, nw = w / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nh * nw) return;
  const int r = i / nw, c = i - r * nw;
  const double* p = fm + (int64_t)(2 * r) * w + 2 * c;
  out[i] = fmax(fmax(p[0], p[1]), fmax(p[w], p[w + 1]));
}

}  // namespace pb
