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
struct StepBookkeeping {   // per-step housekeeping folded into the softmax kernel (block 0) to save graph nodes
  double* norm2;           // per-tensor squared-norm accumulators to zero for this step's grad_sqnorm_kernel
  int n_norm;
  int64_t* cursor;         // cursor[0] += advance (next batch of the permutation); cursor[1] += 1 (step count t)
  int64_t advance;
};

__global__ void advance_cursor_kernel(int64_t* cursor, int64_t advance) {   // tail of the prefetch branch
  if (threadIdx.x == 0) cursor[0] += advance;
}

__global__ void step_bookkeeping_kernel(StepBookkeeping bk) {   // ranks whose shard of a batch is empty
  pdl_wait();
  if (bk.norm2 && (int)threadIdx.x < bk.n_norm) bk.norm2[threadIdx.x] = 0.0;
  if (bk.cursor && threadIdx.x == 0) { bk.cursor[0] += bk.advance; bk.cursor[1] += 1; }
}

__global__ void __launch_bounds__(256) softmax_ce_kernel(const float* __restrict__ logits, int bs, int C,
                                                         int64_t ld, const int32_t* __restrict__ labels,
                                                         float* __restrict__ probs, float* __restrict__ dz,
                                                         int32_t* __restrict__ pred, float* __restrict__ conf,
                                                         double* __restrict__ stats, float* __restrict__ db,
                                                         StepBookkeeping bk) {
  extern __shared__ float s_db[];   // [C] when db != nullptr: block-level bias-gradient partial sums
  pdl_wait();
  pdl_launch_dependents();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (blockIdx.x == 0) {
    if (bk.norm2 && threadIdx.x < bk.n_norm) bk.norm2[threadIdx.x] = 0.0;
    if (bk.cursor && threadIdx.x == 0) { bk.cursor[0] += bk.advance; bk.cursor[1] += 1; }
  }
  if (db) {
    for (int j = threadIdx.x; j < C; j += blockDim.x) s_db[j] = 0.f;
    __syncthreads();
  }
  double my_loss = 0.0;
  unsigned long long my_correct = 0ull;
  if (row < bs) {
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
      if (db && j < C) atomicAdd(&s_db[j], d);           // db_o = sum_rows dZ (backward_pass.pyx:72)
      if (j == y) my_loss = -log((double)p + 1e-9);
    }
    if (lane == 0) {
      if (pred) pred[row] = arg;
      if (conf) conf[row] = inv;  // p[argmax] = exp(0)/s
      if (y >= 0 && arg == y) my_correct = 1ull;
    }
  }
  if (db) {
    __syncthreads();
    for (int j = threadIdx.x; j < C; j += blockDim.x) atomicAdd(&db[j], s_db[j]);
  }
  if (stats && labels) {
    // block-level reduction -> one atomic pair per block
    __shared__ double s_loss[8];
    __shared__ unsigned long long s_corr[8];
    my_loss = warp_sum_d(my_loss);
    const int w = threadIdx.x >> 5;
    if (lane == 0) { s_loss[w] = my_loss; s_corr[w] = my_correct; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double l = 0.0;
      unsigned long long c = 0ull;
      for (int i = 0; i < warps_per_block; ++i) { l += s_loss[i]; c += s_corr[i]; }
      atomicAdd(&stats[0], l);
      atomicAdd(&stats[1], (double)c);   // counts stay exact in a double and reduce with ncclFloat64
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K6(c): db[n] = sum_rows dZ[m, n]  (backward_pass.pyx:72,93).  dz [bs, ld].  grid (ceil(N/32), R);
// block (32, 8): thread (x, y) sums rows y, y+8, ... of its row chunk for column x in FP64, the 8 partials
// are combined through smem and one atomicAdd per (column, chunk) lands in db (zeroed by the caller
// when R > 1).  Coalesced 128-byte row segments; R is chosen so the grid fills the SMs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dz, int bs, int N, int64_t ld,
                                                     float* __restrict__ db, int rows_per_block, int accumulate) {
  __shared__ double red[8][33];
  pdl_wait();
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
    if (gridDim.y == 1 && !accumulate) db[n] = (float)s;
    else atomicAdd(&db[n], (float)s);
  }
}

// ---------------------------------------------------------------------------------------------
// K7: per-tensor squared L2 norm of the (all-reduced) gradient, double accumulation.
// gradient_decent.pyx:33 (np.linalg.norm per tensor).  One block per chunk of the flat buffer.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const float* __restrict__ grads,
                                                          const Chunk* __restrict__ chunks,
                                                          double* __restrict__ norm2, unsigned long long* trace) {
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(trace);
  const Chunk ch = chunks[blockIdx.x];
  const float4* g4 = reinterpret_cast<const float4*>(grads + ch.off);
  const int64_t n4 = ch.len >> 2;  // chunk offsets/lengths are multiples of 4
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = g4[i];
    acc = fmaf(v.x, v.x, acc);
    acc = fmaf(v.y, v.y, acc);
    acc = fmaf(v.z, v.z, acc);
    acc = fmaf(v.w, v.w, acc);
  }
  double d = warp_sum_d((double)acc);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i];
    atomicAdd(&norm2[ch.tensor], t);
  }
  trace_end(trace);
}

struct UpdateParams {
  int opt;       // 0 sgd, 1 adam
  int rule;      // 0 cpu (clip + skip), 1 cupy
  float lr, beta1, beta2, eps, inv_bc1, inv_bc2, eps2, max_norm;
  double beta1_d, beta2_d;
  const int64_t* steps_done;  // device cursor; [1] = Adam step count t of the current step (CuPy rule)
};



// clipped SGD / Adam over the flat buffers; each block handles one chunk and reads its tensor's norm.
// CPU rule (gradient_decent.pyx): g *= 5/||g|| when ||g|| > 5 (:33-35); non-finite -> skip tensor
// (:37-39); Adam m,v in place (:91-95), v_hat floored at eps^2 (:98), w -= lr*m_hat/(sqrt(v_hat)+eps).
__global__ void __launch_bounds__(256) clip_update_kernel(float* __restrict__ params,
                                                          const float* __restrict__ grads,
                                                          float* __restrict__ m, float* __restrict__ v,
                                                          const Chunk* __restrict__ chunks,
                                                          const double* __restrict__ norm2, UpdateParams u,
                                                          unsigned long long* trace) {
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
  if (u.rule == 0) {
    const double n2 = norm2[ch.tensor];
    if (!isfinite(n2)) return;  // NaN/Inf gradient: skip this tensor's update
    const double nrm = sqrt(n2);
    if (nrm > (double)u.max_norm) scale = (float)((double)u.max_norm / nrm);
  }
  const int64_t n4 = ch.len >> 2;
  float4* p4 = reinterpret_cast<float4*>(params + ch.off);
  const float4* g4 = reinterpret_cast<const float4*>(grads + ch.off);
  if (u.opt == 0) {
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p = p4[i];
      const float4 g = g4[i];
      p.x -= u.lr * (g.x * scale);
      p.y -= u.lr * (g.y * scale);
      p.z -= u.lr * (g.z * scale);
      p.w -= u.lr * (g.w * scale);
      p4[i] = p;
    }
    trace_end(trace);
    return;
  }
  float4* m4 = reinterpret_cast<float4*>(m + ch.off);
  float4* v4 = reinterpret_cast<float4*>(v + ch.off);
  const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
  for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p = p4[i], mm = m4[i], vv = v4[i];
    const float4 g = g4[i];
    float pe[4] = {p.x, p.y, p.z, p.w}, me[4] = {mm.x, mm.y, mm.z, mm.w}, ve[4] = {vv.x, vv.y, vv.z, vv.w};
    const float ge[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      me[k] = me[k] * u.beta1 + omb1 * ge[k];
      ve[k] = ve[k] * u.beta2 + omb2 * (ge[k] * ge[k]);
      const float mh = me[k] * u.inv_bc1;
      float vh = ve[k] * u.inv_bc2;
      if (u.rule == 0) vh = fmaxf(vh, u.eps2);
      pe[k] -= u.lr * mh / (sqrtf(vh) + u.eps);
    }
    p4[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    m4[i] = make_float4(me[0], me[1], me[2], me[3]);
    v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
  }
  trace_end(trace);
}

// K7 + update in one launch (CPU rule, networks whose chunks are all co-resident: host checks): every block keeps
// its chunk's gradient in registers, adds its sum of squares to the tensor's norm, meets the other blocks at a
// grid barrier, then clips and updates from the registers -- one launch and one read of the gradient less.
// bar[0] counts block arrivals since the context was created (64-bit, never reset), bar[1] the launches.
__global__ void __launch_bounds__(256, 4) norm_clip_update_kernel(float* __restrict__ params,
                                                               const float* __restrict__ grads,
                                                               float* __restrict__ m, float* __restrict__ v,
                                                               const Chunk* __restrict__ chunks,
                                                               double* __restrict__ norm2, UpdateParams u,
                                                               unsigned long long* __restrict__ bar) {
  pdl_wait();
  const Chunk ch = chunks[blockIdx.x];
  const int64_t n4 = ch.len >> 2;            // <= 512: two 128-bit elements per thread
  const float4* g4 = reinterpret_cast<const float4*>(grads + ch.off);
  float4 g[2];
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int64_t i = threadIdx.x + e * 256;
    g[e] = (i < n4) ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    acc = fmaf(g[e].x, g[e].x, acc); acc = fmaf(g[e].y, g[e].y, acc);
    acc = fmaf(g[e].z, g[e].z, acc); acc = fmaf(g[e].w, g[e].w, acc);
  }
  __shared__ double s[8];
  const double d = warp_sum_d((double)acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(&norm2[ch.tensor], t);
    // grid barrier: every block read bar[1] before it arrives, block 0 bumps it once all have arrived
    const unsigned long long launch = *reinterpret_cast<volatile unsigned long long*>(bar + 1);
    const unsigned long long target = (launch + 1ull) * gridDim.x;
    __threadfence();
    atomicAdd(bar, 1ull);
    unsigned int polls = 0;
    while (*reinterpret_cast<volatile unsigned long long*>(bar) < target) {
      if (++polls > 64u) __nanosleep(32);
      if (polls > (1u << 24)) __trap();   // blocks not co-resident: protocol bug, do not hang
    }
    __threadfence();
    if (blockIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(bar + 1) = launch + 1ull;
  }
  __syncthreads();
  pdl_launch_dependents();                   // only now: a dependent kernel's CTAs must never take slots this grid still needs
  const double n2 = __ldcg(&norm2[ch.tensor]);
  if (!isfinite(n2)) return;                 // NaN/Inf gradient: skip this tensor's update (gradient_decent.pyx:37-39)
  const double nrm = sqrt(n2);
  const float scale = (nrm > (double)u.max_norm) ? (float)((double)u.max_norm / nrm) : 1.0f;
  float4* p4 = reinterpret_cast<float4*>(params + ch.off);
  float4* m4 = reinterpret_cast<float4*>(m + ch.off);
  float4* v4 = reinterpret_cast<float4*>(v + ch.off);
  const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int64_t i = threadIdx.x + e * 256;
    if (i >= n4) continue;
    const float4 p = p4[i];
    float pe[4] = {p.x, p.y, p.z, p.w};
    const float ge[4] = {g[e].x * scale, g[e].y * scale, g[e].z * scale, g[e].w * scale};
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
        const float vh = fmaxf(ve[k] * u.inv_bc2, u.eps2);
        pe[k] -= u.lr * mh / (sqrtf(vh) + u.eps);
      }
      m4[i] = make_float4(me[0], me[1], me[2], me[3]);
      v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    }
    p4[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
  }
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

// split-K fix-up: out = epilogue(acc) for GEMMs whose partial sums were accumulated with atomics.
// epi: 0 none, 1 +bias, 2 relu(+bias), 3 mask by (act > 0)
__global__ void gemm_fixup_kernel(float* __restrict__ c, int M, int N, int64_t ldc, int epi,
                                  const float* __restrict__ bias, const float* __restrict__ mask,
                                  int64_t ldmask) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / N, n = i - r * N;
    float v = c[r * ldc + n];
    if (epi == 1 || epi == 2) v += bias[n];
    if (epi == 2) v = (v < 0.f) ? 0.f : v;
    if (epi == 3) v = (mask[r * ldmask + n] <= 0.f) ? 0.f : v;
    c[r * ldc + n] = v;
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
