// mlp_tail.cuh -- the shallow tail of the MLP forward pass as ONE kernel.
//
// At BASELINE configs[1] the forward pass after layer 1 is three tiny GEMMs (4096x128x256, 4096x64x128,
// 4096x10x64) plus the split-K fix-up of layer 1 and the softmax/CE kernel: five dependent launches of 5-11 us
// each that are pure latency (launch, prologue, TMA->MMA->TMEM->global round trips).  Here a CTA owns 128 batch
// rows and walks the whole tail with the activations never leaving the SM:
//
//   z1 tile (TMA, 128 x n1)  --+b1, ReLU in place-->  A operand in smem (K-major, SWIZZLE_128B)  [also stored: a1]
//   for every tail layer:  tcgen05.mma over the k-blocks (weights streamed through a 2-stage TMA ring)
//                          -> TMEM -> +bias, ReLU -> next layer's A operand written straight into the swizzled
//                          smem layout (and to global for the backward pass)
//   output layer:          logits stay in registers (one thread = one row): softmax, CE, arg-max accuracy,
//                          dZ = p - y, output-bias gradient, per-step bookkeeping  (softmax_ce_kernel's job)
//
// warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = transform/epilogue (two warps per TMEM
// lane quarter; each thread owns a row and every other 32-column block of it).
// STATUS (round 1): correct (same parity tests as the per-layer path) but not yet faster -- 30 us against the
// 36 us of the five launches it replaces at configs[1], because with 32 CTAs each CTA's chain of hand-offs is
// exposed; it is therefore opt-in (PYCNN_B200_FUSED_TAIL=1) until the chain is pipelined across two row tiles.
// Eligible when the layer-1 width is <= 256, every later width <= 128 and C <= 64; wider nets (configs 3, 5) keep the per-layer GEMMs, which are
// large enough there to run at 60+ % of the tensor peak on their own.
// Reference semantics: forward_pass.pyx:28-41 (ReLU as x<0 -> 0), optimized.cpp:34-46, pycnn.py:634-639,
// backward_pass.pyx:65,72.
#pragma once
#include "elementwise.cuh"
#include "tc_common.cuh"

namespace pb {
namespace tail {

constexpr int MAX_TAIL = 4;       // tail layers incl. the output layer
constexpr int MAX_W = 256;        // widest tail INPUT (layer-1 width) handled on chip
constexpr int MAX_N = 128;        // widest tail layer OUTPUT (one TMA box of <= 128 weight rows per k-block)
constexpr int MAX_C = 64;
constexpr int ACT_BYTES = 128 * MAX_W * 4;      // 128 KB: one 128-row activation tile
constexpr int W_STAGE_BYTES = MAX_N * 32 * 4;   // 16 KB: one 32-wide k-block of up to 128 output rows
constexpr int W_STAGES = 5;                     // deep enough to hide the L2 latency of the weight stream
constexpr int EPI_THREADS = 256;                // 8 transform/epilogue warps: two per TMEM lane quarter
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int SMEM_BYTES = ACT_BYTES + W_STAGES * W_STAGE_BYTES + (MAX_TAIL + 1) * MAX_W * 4 + 256;

struct TailArgs {
  int bs, nl;
  int K[MAX_TAIL];           // input width of tail layer j
  int N[MAX_TAIL];           // output width of tail layer j (last = C)
  int fuse_fixup;            // input tile is raw z1: apply +bias0, ReLU, and store a1
  const float* bias0;        // b1
  float* a_in;               // [bs, ld_in]: a1 is written here when fuse_fixup
  int64_t ld_in;
  const float* bias[MAX_TAIL];
  float* act_out[MAX_TAIL];  // hidden outputs (j < nl-1); unused for the last layer
  int64_t ld_out[MAX_TAIL];
  // softmax / CE / dZ
  const int32_t* labels;
  float* dz;
  int64_t ld_dz;
  double* stats;
  float* db_out;
  StepBookkeeping bk;
};

// byte offset of (row r, column k) inside a K-major SWIZZLE_128B activation tile of 32-wide k-blocks
__device__ __forceinline__ uint32_t act_off(int r, int k) {
  const int kb = k >> 5, c = (k >> 2) & 7;
  return (uint32_t)(kb * 16384 + r * 128 + ((c ^ (r & 7)) << 4) + ((k & 3) << 2));
}

template <int CPAD>   // class columns kept in registers by the softmax stage: 16, 32 or 64 (>= C)
__global__ void __launch_bounds__(THREADS, 1)
mlp_tail_forward_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_w0,
                        const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2,
                        const __grid_constant__ CUtensorMap map_w3, const __grid_constant__ TailArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* act = smem;                                   // [k-block][128 rows][128 B] swizzled
  uint8_t* wring = smem + ACT_BYTES;                     // W_STAGES x [rows][128 B] swizzled
  float* s_bias = reinterpret_cast<float*>(wring + W_STAGES * W_STAGE_BYTES);   // [MAX_TAIL+1][MAX_W]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + (MAX_TAIL + 1) * MAX_W);
  uint64_t* act_full = bars;          // producer -> transform (tx)
  uint64_t* act_ready = bars + 1;     // transform/epilogue -> MMA (128 arrivals per layer)
  uint64_t* acc_full = bars + 2;      // MMA -> epilogue (commit, once per layer)
  uint64_t* w_full = bars + 3;        // [W_STAGES]
  uint64_t* w_empty = w_full + W_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + W_STAGES);
  __shared__ float s_db[MAX_C];
  __shared__ double s_loss[4];
  __shared__ double s_corr[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const CUtensorMap* wmaps[MAX_TAIL] = {&map_w0, &map_w1, &map_w2, &map_w3};

  if (threadIdx.x == 0) {
    tc::prefetch_tensormap(&map_in);
    for (int j = 0; j < a.nl; ++j) tc::prefetch_tensormap(wmaps[j]);
    tc::mbar_init(act_full, 1);
    tc::mbar_init(act_ready, EPI_THREADS);
    tc::mbar_init(acc_full, 1);
    for (int s = 0; s < W_STAGES; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
  if (threadIdx.x < MAX_C) s_db[threadIdx.x] = 0.f;
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  // Parameters (biases, weights) were last written by the PREVIOUS step's update kernel, which finished before the
  // kernel we depend on even started: they may be read before griddepcontrol.wait, overlapping layer 1's tail.
  for (int i = threadIdx.x; i < (a.nl + 1) * MAX_W; i += blockDim.x) {
    const int j = i / MAX_W, k = i - j * MAX_W;
    float v = 0.f;
    if (j == 0) { if (a.fuse_fixup && k < a.K[0]) v = a.bias0[k]; }
    else if (k < a.N[j - 1]) v = a.bias[j - 1][k];
    s_bias[i] = v;
  }
  __syncthreads();

  if (warp == 0) {
    // ===== TMA producer: the input tile once, then every layer's weight k-blocks through the ring =====
    if (lane == 0) {
      const int nkb0 = (a.K[0] + 31) >> 5;
      bool input_issued = false;
      int it = 0;
      for (int j = 0; j < a.nl; ++j) {
        const int nkb = (a.K[j] + 31) >> 5;
        const int rows = (a.N[j] + 15) & ~15;          // MMA N (multiple of 16); TMA box rows of this map
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % W_STAGES;
          if (!input_issued && it == W_STAGES) {
            // the weight ring is full (all of it fetched while the previous kernel was still draining):
            // now wait for z1 and fetch the input tile
            pdl_wait();
            tc::mbar_arrive_expect_tx(act_full, (uint32_t)(nkb0 * 16384));
            for (int k2 = 0; k2 < nkb0; ++k2) tc::tma_load_2d(act + k2 * 16384, &map_in, act_full, k2 * 32, m0);
            input_issued = true;
          }
          tc::mbar_wait(&w_empty[s], ((it / W_STAGES) & 1) ^ 1);
          tc::mbar_arrive_expect_tx(&w_full[s], (uint32_t)(rows * 128));
          tc::tma_load_2d(wring + s * W_STAGE_BYTES, wmaps[j], &w_full[s], kb * 32, 0);
        }
      }
      if (!input_issued) {   // fewer weight k-blocks than ring stages
        pdl_wait();
        tc::mbar_arrive_expect_tx(act_full, (uint32_t)(nkb0 * 16384));
        for (int k2 = 0; k2 < nkb0; ++k2) tc::tma_load_2d(act + k2 * 16384, &map_in, act_full, k2 * 32, m0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < a.nl; ++j) {
        const int nkb = (a.K[j] + 31) >> 5;
        const int npad = (a.N[j] + 15) & ~15;
        const uint32_t idesc = tc::make_idesc(tc::FMT_TF32, 128, npad, false, false);
        tc::mbar_wait(act_ready, (uint32_t)(j & 1));   // A operand of layer j is in smem (and TMEM is drained)
        tc::fence_after_thread_sync();
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % W_STAGES;
          tc::mbar_wait(&w_full[s], (it / W_STAGES) & 1);
          tc::fence_after_thread_sync();
          const uint32_t sa = tc::smem_u32(act + kb * 16384);
          const uint32_t sb = tc::smem_u32(wring + s * W_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = tc::make_smem_desc(sa + k * 32, 16, 1024, tc::UMMA_SWIZZLE_128B);
            const uint64_t db = tc::make_smem_desc(sb + k * 32, 16, 1024, tc::UMMA_SWIZZLE_128B);
            tc::mma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc::mma_commit(&w_empty[s]);
        }
        tc::mma_commit(acc_full);
      }
    }
  } else {
    // ===== transform / epilogue warps: thread = batch row =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;            // 0/1: which alternate k-blocks of the row this thread handles
    const int r = q * 32 + lane;                 // row inside the tile == TMEM lane
    const int row = m0 + r;
    const bool row_ok = row < a.bs;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    // ---- phase 0: the input tile
    pdl_wait();                       // z1 (and everything before it) is complete before a1 is stored over it
    pdl_launch_dependents();
    tc::mbar_wait(act_full, 0);
    if (a.fuse_fixup) {
      const int nkb0 = (a.K[0] + 31) >> 5;
      float* grow = a.a_in + (int64_t)row * a.ld_in;
      for (int kb = half; kb < nkb0; kb += 2) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int k = kb * 32 + c * 4;
          float4* p = reinterpret_cast<float4*>(act + act_off(r, k));
          float4 v = *p;
          const float4 b = *reinterpret_cast<const float4*>(s_bias + k);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
          v.x = (v.x < 0.f) ? 0.f : v.x; v.y = (v.y < 0.f) ? 0.f : v.y;     // relu_inplace: NaN stays NaN
          v.z = (v.z < 0.f) ? 0.f : v.z; v.w = (v.w < 0.f) ? 0.f : v.w;
          if (k >= a.K[0]) v = make_float4(0.f, 0.f, 0.f, 0.f);             // K is a multiple of 4 (row pitch)
          *p = v;
          if (row_ok && k < a.K[0]) *reinterpret_cast<float4*>(grow + k) = v;   // a1 for the backward pass
        }
      }
      tc::fence_proxy_async_smem();
    }
    tc::mbar_arrive(act_ready);
    // ---- layers
    for (int j = 0; j < a.nl; ++j) {
      const bool last = (j == a.nl - 1);
      const int n = a.N[j];
      const float* bj = s_bias + (j + 1) * MAX_W;
      tc::mbar_wait(acc_full, (uint32_t)(j & 1));
      tc::fence_after_thread_sync();
      if (!last) {
        float* grow = a.act_out[j] + (int64_t)row * a.ld_out[j];
        const int nkb_next = (n + 31) >> 5;
        for (int kb = half; kb < nkb_next; kb += 2) {
          float v[32];
          __syncwarp();
          tc::tmem_ld32(taddr + kb * 32, v);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k = kb * 32 + c * 4;
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = v[c * 4 + e] + bj[k + e];
              x = (x < 0.f) ? 0.f : x;
              o[e] = (k + e < n) ? x : 0.f;           // columns beyond the layer width feed zeros to the next MMA
            }
            const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(act + act_off(r, k)) = o4;
            if (row_ok && k < n) {
              if (k + 3 < n) *reinterpret_cast<float4*>(grow + k) = o4;
              else for (int e = 0; e < 4; ++e) if (k + e < n) grow[k + e] = o[e];
            }
          }
        }
        tc::fence_proxy_async_smem();       // next layer's A operand is visible to the tensor core
        tc::fence_before_thread_sync();     // this layer's tcgen05.ld are done before the next MMA overwrites TMEM
        tc::mbar_arrive(act_ready);
      } else if (half == 0) {
        // ---- output layer (warps 2-5 only): logits in registers -> softmax, CE, arg-max, dZ, db_o
        const int C = n;
        float z[CPAD];
        if constexpr (CPAD == 16) {
          float v[16];
          __syncwarp();
          tc::tmem_ld16(taddr, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) z[e] = v[e] + bj[e];
        } else {
#pragma unroll
          for (int cb = 0; cb < CPAD / 32; ++cb) {
            float v[32];
            __syncwarp();
            tc::tmem_ld32(taddr + cb * 32, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) z[cb * 32 + e] = v[e] + bj[cb * 32 + e];
          }
        }
        float mx = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < CPAD; ++k)
          if (k < C && z[k] > mx) { mx = z[k]; arg = k; }        // first maximum: np.argmax semantics
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < CPAD; ++k)
          if (k < C) { z[k] = expf(z[k] - mx); s += z[k]; }
        const float inv = 1.0f / s;
        const int y = row_ok ? a.labels[row] : -1;
        double my_loss = 0.0, my_corr = 0.0;
        float* dzrow = a.dz + (int64_t)row * a.ld_dz;
        float py = 1.0f;
#pragma unroll
        for (int k = 0; k < CPAD; ++k) {
          if (k < (int)a.ld_dz) {
            const float p = (k < C) ? z[k] * inv : 0.f;
            const float d = (k < C) ? (p - (k == y ? 1.0f : 0.0f)) : 0.f;
            if (k == y) py = p;                  // a select, not a branch: the (double) log below is emitted once
            if (row_ok) dzrow[k] = d;
            z[k] = row_ok ? d : 0.f;
          }
        }
        if (row_ok) my_loss = -log((double)py + 1e-9);
        if (row_ok && arg == y) my_corr = 1.0;
        // db_o: column sums over the CTA's rows (warp shuffles, then one shared atomic per warp and column)
#pragma unroll
        for (int k = 0; k < CPAD; ++k) {
          if (k < C) {
            float d = z[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            if (lane == 0) atomicAdd(&s_db[k], d);
          }
        }
        my_loss = warp_sum_d(my_loss);
        my_corr = warp_sum_d(my_corr);
        if (lane == 0) { s_loss[q] = my_loss; s_corr[q] = my_corr; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = threadIdx.x - 64;     // 0..127 among the epilogue threads
        if (t < C) atomicAdd(&a.db_out[t], s_db[t]);
        if (t == 0) {
          atomicAdd(&a.stats[0], s_loss[0] + s_loss[1] + s_loss[2] + s_loss[3]);
          atomicAdd(&a.stats[1], s_corr[0] + s_corr[1] + s_corr[2] + s_corr[3]);
          if (blockIdx.x == 0 && a.bk.cursor) { a.bk.cursor[0] += a.bk.advance; a.bk.cursor[1] += 1; }
        }
        if (blockIdx.x == 0 && a.bk.norm2 && t < a.bk.n_norm) a.bk.norm2[t] = 0.0;
      }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace tail
}  // namespace pb
