// tail_fused.cuh -- everything between the two layer-1 GEMMs of a training step as ONE kernel.
//
// At BASELINE configs[1] (4096 x [4275, 256, 128, 64, 10]) the forward pass after layer 1 and the whole dA chain of
// the backward pass are six GEMMs of 0.03-0.27 GFLOP each.  As separate launches they cost 5-8 us apiece -- launch
// hand-off, a cold TMA round trip, an SM-ingest-bound mainloop that re-reads the activation tile once per N tile and
// an epilogue through global memory -- 42 us of a 113 us step for 5 % of its FLOPs.  Here a CTA owns 128 batch rows
// and walks the chain with the activations never leaving the SM:
//
//   forward  a_l   = relu(a_{l-1} . W_l^T + b_l)            l = 2..L     (forward_pass.pyx:28-34)
//   output   p     = softmax(a_L . W_o^T + b_o), CE, arg-max, dZ_o = p - y, db_o        (:36-41, optimized.cpp:34-46,
//                                                                        pycnn.py:634-640, backward_pass.pyx:65,72)
//   backward dZ_{l-1} = (dZ_l . W_l) masked by a_{l-1} > 0, db_{l-1} = column sums      (backward_pass.pyx:85-97)
//
// Every stage is one tcgen05 GEMM  D[128 x N] = A[128 x K] . B:  A is the previous stage's output, written by the
// epilogue straight into the K-major SWIZZLE_128B operand layout in shared memory (only the first stage streams its A,
// the layer-1 activation tile, from global memory k-block by k-block); B (the weights: K-major for the forward stages,
// MN-major for the dA stages) streams through a 3-slot TMA ring that the producer fills ahead of the chain, so the
// weight latency hides behind the previous stage's epilogue; D lives in TMEM.  Activations / dZ also go to global
// memory once, for the dW GEMMs that follow.  ReLU masks never come back from memory: an epilogue thread owns a row
// end to end and keeps the sign bits of what it produced.  Bias gradients leave as one ordered partial per row tile
// (GradParts), losses as integer atomics: bit-reproducible like the per-layer path, and bit-IDENTICAL to it for the
// activations (same k order, same rounding).
//
// warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue (thread = TMEM lane = batch row).
// Eligible (host): TF32 mode, 1 <= L <= 4, layer-1 width <= 256, later widths <= 128, C <= 32.
#pragma once
#include "elementwise.cuh"
#include "tc_common.cuh"

namespace pb {
namespace tail {

constexpr int MAX_STAGES = 8;                // 2 L stages
constexpr int MAX_IN = 256;                  // widest stage-0 input / last-stage output (layer-1 width)
constexpr int MAX_MID = 128;                 // widest activation kept on chip
constexpr int SLOT_BYTES = 32 * 1024;        // ring slot: A k-block (16 KB, first stage only) + B k-block (<= 16 KB)
constexpr int RING = 3;
constexpr int ACT_BYTES = 128 * MAX_MID * 4; // 64 KB: one on-chip activation tile (4 k-blocks of [128 rows][128 B])
constexpr int THREADS = 192;
constexpr int SMEM_BYTES = RING * SLOT_BYTES + 2 * ACT_BYTES + 256 /*barriers*/;

struct Stage {
  int K, N;              // contraction length, output width
  int b_mn;              // 0: B = W [N x K] row-major (forward: K-major), 1: B = W [K x N] row-major (dA: MN-major)
  int epi;               // 0: +bias, ReLU (hidden forward), 1: softmax / CE / dZ (output layer), 2: ReLU mask (backward)
  int mask_slot;         // epi 0: slot the produced sign bits go to; epi 2: slot whose bits mask this output
  const float* bias;     // epi 0, 1
  float* out;            // [bs, ld_out] row-major: activation (epi 0) or dZ (epi 1, 2); never null
  int64_t ld_out;
  float* colsum_part;    // epi 1, 2: this row tile's bias-gradient partial at colsum_part[tile * ld_colsum + n]
  int64_t ld_colsum;
};

struct TailArgs {
  int bs, nstages;
  Stage st[MAX_STAGES];
  const float* a_in;     // layer-1 activations [bs, ld_in] (stage 0's A; also the mask of the last stage)
  int64_t ld_in;
  int in_width;          // layer-1 width
  const int32_t* labels;
  unsigned long long* stats;
  StepBookkeeping bk;
  unsigned long long* trace;
};

// byte offset of (row r, column k) inside a K-major SWIZZLE_128B activation tile of 32-wide k-blocks
__device__ __forceinline__ uint32_t act_off(int r, int k) {
  const int kb = k >> 5, c = (k >> 2) & 7;
  return (uint32_t)(kb * 16384 + r * 128 + ((c ^ (r & 7)) << 4) + ((k & 3) << 2));
}

// 32 values per lane (one row each) -> lane j ends with the sum of column j over the warp's 32 rows: 31 shuffles,
// a fixed summation tree
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = hi ? v[i] : v[i + o];
      const float keep = hi ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_b0,
                  const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                  const __grid_constant__ CUtensorMap map_b3, const __grid_constant__ CUtensorMap map_b4,
                  const __grid_constant__ CUtensorMap map_b5, const __grid_constant__ CUtensorMap map_b6,
                  const __grid_constant__ CUtensorMap map_b7, const __grid_constant__ TailArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* ring = smem;
  uint8_t* act[2] = {smem + RING * SLOT_BYTES, smem + RING * SLOT_BYTES + ACT_BYTES};
  // column-sum staging [4 warps][MAX_IN] lives in slot 0's A region: only stage 0 streams an A operand, and a stage's
  // column sums are written after all of its MMAs have completed
  float* cs_stage = reinterpret_cast<float*>(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING * SLOT_BYTES + 2 * ACT_BYTES);
  uint64_t* full_bar = bars;                 // [RING]
  uint64_t* empty_bar = bars + RING;         // [RING]
  uint64_t* acc_full = bars + 2 * RING;      // MMA -> epilogue, once per stage
  uint64_t* act_ready = acc_full + 1;        // epilogue -> MMA, 128 arrivals per stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const CUtensorMap* bmaps[MAX_STAGES] = {&map_b0, &map_b1, &map_b2, &map_b3, &map_b4, &map_b5, &map_b6, &map_b7};

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_in);
    for (int s = 0; s < a.nstages; ++s) prefetch_tensormap(bmaps[s]);
    for (int s = 0; s < RING; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(act_ready, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(a.trace);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: stage 0's A k-blocks and every stage's B k-blocks, in the order the MMA issuer consumes them
      int it = 0;
      for (int s = 0; s < a.nstages; ++s) {
        const Stage& S = a.st[s];
        const int nkb = (S.K + 31) >> 5;
        const int halves = (S.N + 127) >> 7;
        for (int h = 0; h < halves; ++h) {
          const int nh = min(128, S.N - h * 128);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int slot = it % RING;
            mbar_wait(&empty_bar[slot], ((it / RING) & 1) ^ 1);
            uint8_t* sa = ring + slot * SLOT_BYTES;
            uint8_t* sb = sa + 16384;
            if (S.b_mn) {
              const int atoms = (nh + 31) >> 5;
              mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)atoms * 4096u + (s == 0 ? 16384u : 0u));
              for (int at = 0; at < atoms; ++at) tma_load_2d(sb + at * 4096, bmaps[s], &full_bar[slot], h * 128 + at * 32, kb * 32);
            } else {
              const int rows = (nh + 15) & ~15;          // box rows of the K-major weight map
              mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)rows * 128u + (s == 0 ? 16384u : 0u));
              tma_load_2d(sb, bmaps[s], &full_bar[slot], kb * 32, h * 128);
            }
            if (s == 0) tma_load_2d(sa, &map_in, &full_bar[slot], kb * 32, m0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      int it = 0;
      for (int s = 0; s < a.nstages; ++s) {
        const Stage& S = a.st[s];
        const int nkb = (S.K + 31) >> 5;
        const int halves = (S.N + 127) >> 7;
        if (s > 0) {                                      // A = the previous stage's output, just written by the epilogue
          mbar_wait(act_ready, (s - 1) & 1);
          fence_after_thread_sync();
        }
        const uint32_t a_tile = smem_u32(act[(s - 1) & 1]);
        for (int h = 0; h < halves; ++h) {
          const int nh = min(128, S.N - h * 128);
          const int nr = S.b_mn ? ((nh + 31) & ~31) : ((nh + 15) & ~15);
          const uint32_t idesc = make_idesc(FMT_TF32, 128, nr, false, S.b_mn != 0);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int slot = it % RING;
            mbar_wait(&full_bar[slot], (it / RING) & 1);
            fence_after_thread_sync();
            const uint32_t sa = (s == 0) ? smem_u32(ring + slot * SLOT_BYTES) : a_tile + kb * 16384;
            const uint32_t sb = smem_u32(ring + slot * SLOT_BYTES + 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
              const uint64_t db = S.b_mn ? make_smem_desc(sb + k * 1024, 4096, 512, UMMA_SWIZZLE_128B_BASE32B)
                                         : make_smem_desc(sb + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
              mma_tf32(tmem_base + h * 128, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&empty_bar[slot]);
          }
        }
        mma_commit(acc_full);
      }
    }
  } else {
    // ===== epilogue warps: thread = accumulator row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t grow = (int64_t)m0 + row;
    const bool valid = grow < a.bs;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t mask_bits[4][8];                 // [slot][32-column word]: slot 0 = layer-1 activations, 1.. = produced here
    // sign bits of this row of the layer-1 activations (mask of the LAST stage), read while stage 0's mainloop runs
#pragma unroll 1
    for (int w = 0; w < 8; ++w) {
      uint32_t bits = 0u;
      if (valid && w * 32 < a.in_width) {
        const float4* src = reinterpret_cast<const float4*>(a.a_in + grow * a.ld_in + w * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (w * 32 + 4 * j < a.in_width) {          // rows are padded to 4 floats with zeros
            const float4 x = __ldg(src + j);
            bits |= (x.x > 0.f ? 1u : 0u) << (4 * j) | (x.y > 0.f ? 1u : 0u) << (4 * j + 1) |
                    (x.z > 0.f ? 1u : 0u) << (4 * j + 2) | (x.w > 0.f ? 1u : 0u) << (4 * j + 3);
          }
        }
      }
      mask_bits[0][w] = bits;
    }
    if (blockIdx.x == 0 && threadIdx.x == 64) step_bookkeeping(a.bk);
#pragma unroll 1
    for (int s = 0; s < a.nstages; ++s) {
      const Stage& S = a.st[s];
      const bool last = (s == a.nstages - 1);
      uint8_t* nxt = act[s & 1];
      mbar_wait(acc_full, s & 1);
      fence_after_thread_sync();
      if (threadIdx.x == 64) trace_mark(a.trace, s < 3 ? s + 1 : 3);
      if (S.epi == 1) {
        // ---- output layer: logits -> softmax -> CE, arg-max, dZ = p - y, db_o (one 32-column chunk)
        float v[32];
        tmem_ld32(t_lane, v);
        const int C = S.N;
        float mx = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < C) {
            v[j] += __ldg(S.bias + j);
            if (v[j] > mx) { mx = v[j]; arg = j; }     // first maximum: np.argmax tie-break
          }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < C) { v[j] = expf(v[j] - mx); sum += v[j]; }
        const float inv = 1.0f / sum;
        const int y = valid ? a.labels[grow] : -1;
        float py = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float d = 0.f;
          if (j < C && valid) {
            const float p = v[j] * inv;
            if (j == y) py = p;
            d = p - (j == y ? 1.0f : 0.0f);
          }
          v[j] = d;
        }
        if (valid) {
          float4* drow = reinterpret_cast<float4*>(S.out + grow * S.ld_out);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < (int)S.ld_out) drow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (!last) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(nxt + act_off(row, 4 * j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        double loss = valid ? -log((double)py + 1e-9) : 0.0;
        float correct = (valid && arg == y) ? 1.f : 0.f;
        loss = warp_sum_d(loss);
        correct = warp_sum(correct);
        if (lane == 0) stats_add(a.stats, loss, (unsigned long long)correct);
        const float cs = warp_colsum32(v, lane);
        cs_stage[q * MAX_IN + lane] = cs;
      } else {
        const int nchunks = (S.N + 31) >> 5;
#pragma unroll 1
        for (int ch = 0; ch < nchunks; ++ch) {
          const int c = ch * 32;
          float v[32];
          tmem_ld32(t_lane + c, v);
          if (S.epi == 0) {
            uint32_t bits = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = 0.f;
              if (c + j < S.N) {
                x = v[j] + __ldg(S.bias + c + j);
                x = (x < 0.f) ? 0.f : x;               // the reference's comparison: NaN propagates (forward_pass.pyx:8-13)
              }
              v[j] = x;
              bits |= (x > 0.f ? 1u : 0u) << j;
            }
            mask_bits[S.mask_slot & 3][ch & 7] = bits;
          } else {
            const uint32_t bits = mask_bits[S.mask_slot & 3][ch & 7];
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (c + j < S.N && ((bits >> j) & 1u) && valid) ? v[j] : 0.f;     // z <= 0 -> 0 (backward_pass.pyx:86)
          }
          if (!last) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(nxt + act_off(row, c + 4 * j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (valid) {
            float4* orow = reinterpret_cast<float4*>(S.out + grow * S.ld_out + c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c + 4 * j < (int)S.ld_out) orow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (S.epi == 2) {
            const float cs = warp_colsum32(v, lane);
            cs_stage[q * MAX_IN + c + lane] = cs;
          }
        }
      }
      // the next stage's A operand is complete (generic-proxy writes -> async proxy), TMEM is drained
      fence_proxy_async_smem();
      fence_before_thread_sync();
      mbar_arrive(act_ready);
      if (S.epi != 0) {
        // bias gradient of this row tile: the four warps' column sums, added in warp order, one ordered partial
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = (int)threadIdx.x - 64;
        for (int n = t; n < (int)S.ld_colsum; n += 128) {
          const float tot = (n < S.N) ? ((cs_stage[n] + cs_stage[MAX_IN + n]) + (cs_stage[2 * MAX_IN + n] + cs_stage[3 * MAX_IN + n])) : 0.f;
          S.colsum_part[(int64_t)blockIdx.x * S.ld_colsum + n] = tot;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");          // cs_stage is reused by the next stage
      }
    }
  }
  fence_before_thread_sync();
  __syncthreads();
  trace_end(a.trace);
  if (warp == 1) {
    fence_after_thread_sync();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace tail
}  // namespace pb
