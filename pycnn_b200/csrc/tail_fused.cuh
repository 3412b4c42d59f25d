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
constexpr int B_SLOT = 16 * 1024;            // weight ring slot (stages >= 1): one k-block of <= 128 output rows / columns
constexpr int B_RING = 3;
constexpr int S0_SLOT = 32 * 1024;           // stage-0 ring slot: A k-block (16 KB) + B k-block (<= 16 KB)
constexpr int S0_RING = 5;                   // 160 KB pool; its first 128 KB become the two on-chip activation tiles afterwards
constexpr int ACT_BYTES = 128 * MAX_MID * 4; // 64 KB: one activation tile (4 k-blocks of [128 rows][128 B], SWIZZLE_128B)
constexpr int EPI_THREADS = 256;             // 8 epilogue warps: two per TMEM lane quarter, alternating 32-column chunks
constexpr int THREADS = 64 + EPI_THREADS;
constexpr int MASK_WORDS = MAX_IN / 32 + 3 * (MAX_MID / 32);     // sign bits per row: layer-1 width + up to 3 produced layers
constexpr int SMEM_BYTES = B_RING * B_SLOT + S0_RING * S0_SLOT + 256 /*barriers*/ + MAX_STAGES * MAX_MID * 4 /*biases*/ +
                           MASK_WORDS * 128 * 4 /*ReLU sign bits*/;

struct TailMaps {        // one kernel parameter: the maps are picked by stage index with plain address arithmetic
  CUtensorMap in;                    // layer-1 activations (stage 0's A operand)
  CUtensorMap b[MAX_STAGES];         // weights of each stage
  CUtensorMap o[MAX_STAGES];         // output matrix of each stage (TMA store)
};

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

// No thread may touch LOCAL memory or issue global LOADS on the chain: an SM returns loads in issue order, and behind the
// row-gather prefetch sharing the SMs (hundreds of DRAM-latency loads in flight) every such access cost about a
// microsecond -- per-stage arrays are therefore indexed by address arithmetic (parameter space, shared memory) only.
// 96 registers per thread: 320 x 96 leaves room in the register file for the two 256-thread CTAs of the row-gather
// prefetch that runs beside this kernel (2 x 256 x 64); at the compiler's own 123 the gather lost these 32 SMs, ran 39 us
// instead of 30 and delayed the layer-1 dW GEMM behind it.
__global__ void __maxnreg__(96)
tail_fused_kernel(const __grid_constant__ TailMaps maps, const __grid_constant__ TailArgs a) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* bring = smem;                                   // weight ring of stages >= 1
  uint8_t* pool = smem + B_RING * B_SLOT;                  // stage-0 ring, then: act[0] | act[1] | column-sum staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(pool + S0_RING * S0_SLOT);
  uint64_t* full0 = bars;                     // [S0_RING]
  uint64_t* empty0 = full0 + S0_RING;         // [S0_RING]  MMA commit + the 128 epilogue threads (they read a1's signs)
  uint64_t* fullb = empty0 + S0_RING;         // [B_RING]
  uint64_t* emptyb = fullb + B_RING;          // [B_RING]
  uint64_t* acc_full = emptyb + B_RING;       // MMA -> epilogue, once per stage
  uint64_t* act_ready = acc_full + 1;         // epilogue -> MMA, 128 arrivals per stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);
  // Biases of the forward stages, staged once: the epilogue chain must not issue global LOADS -- an SM returns them in
  // issue order, so behind the row-gather prefetch that shares the SMs (hundreds of DRAM-latency loads in flight) every
  // bias fetch cost microseconds per 32-column chunk (measured: 5-8 us per stage; with the staging ~1.5 us)
  float* s_bias = reinterpret_cast<float*>(pool + S0_RING * S0_SLOT + 256);        // [MAX_STAGES][MAX_MID]
  // ReLU sign bits, [word][row]: words 0..7 = the layer-1 activations, then 4 words per layer produced here
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_bias + MAX_STAGES * MAX_MID);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nkb0 = (a.st[0].K + 31) >> 5;     // k-blocks of stage 0 (its N is <= 128: a single half)

  if (threadIdx.x == 0) {
    prefetch_tensormap(&maps.in);
    for (int s = 0; s < a.nstages; ++s) { prefetch_tensormap(&maps.b[s]); prefetch_tensormap(&maps.o[s]); }
    for (int s = 0; s < S0_RING; ++s) { mbar_init(&full0[s], 1); mbar_init(&empty0[s], 129); }
    for (int s = 0; s < B_RING; ++s) { mbar_init(&fullb[s], 1); mbar_init(&emptyb[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(act_ready, EPI_THREADS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  // parameters were last written by the previous step's update, long complete: readable before griddepcontrol.wait
  for (int i = threadIdx.x; i < a.nstages * MAX_MID; i += blockDim.x) {
    const int s = i / MAX_MID, n = i - s * MAX_MID;
    s_bias[i] = (a.st[s].epi != 2 && n < a.st[s].N) ? a.st[s].bias[n] : 0.f;
  }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(a.trace);

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: stage 0's (A, B) k-blocks through the pool, then every later stage's B k-blocks through the
      // weight ring, in the order the MMA issuer consumes them (it runs ahead of the chain by the ring depth)
      {
        const Stage& S = a.st[0];
        const int rows = (min(128, S.N) + 15) & ~15;
        for (int kb = 0; kb < nkb0; ++kb) {
          const int slot = kb % S0_RING;
          mbar_wait(&empty0[slot], ((kb / S0_RING) & 1) ^ 1);
          uint8_t* sa = pool + slot * S0_SLOT;
          mbar_arrive_expect_tx(&full0[slot], 16384u + (uint32_t)rows * 128u);
          tma_load_2d(sa, &maps.in, &full0[slot], kb * 32, m0);
          tma_load_2d(sa + 16384, &maps.b[0], &full0[slot], kb * 32, 0);
        }
      }
      int it = 0;
      for (int s = 1; s < a.nstages; ++s) {
        const Stage& S = a.st[s];
        const int nkb = (S.K + 31) >> 5;
        const int halves = (S.N + 127) >> 7;
        for (int h = 0; h < halves; ++h) {
          const int nh = min(128, S.N - h * 128);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int slot = it % B_RING;
            mbar_wait(&emptyb[slot], ((it / B_RING) & 1) ^ 1);
            uint8_t* sb = bring + slot * B_SLOT;
            if (S.b_mn) {
              const int atoms = (nh + 31) >> 5;
              mbar_arrive_expect_tx(&fullb[slot], (uint32_t)atoms * 4096u);
              for (int at = 0; at < atoms; ++at) tma_load_2d(sb + at * 4096, &maps.b[s], &fullb[slot], h * 128 + at * 32, kb * 32);
            } else {
              const int rows = (nh + 15) & ~15;          // box rows of the K-major weight map
              mbar_arrive_expect_tx(&fullb[slot], (uint32_t)rows * 128u);
              tma_load_2d(sb, &maps.b[s], &fullb[slot], kb * 32, h * 128);
            }
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
        const uint32_t a_tile = smem_u32(pool + ((s - 1) & 1) * ACT_BYTES);
        for (int h = 0; h < halves; ++h) {
          const int nh = min(128, S.N - h * 128);
          const int nr = S.b_mn ? ((nh + 31) & ~31) : ((nh + 15) & ~15);
          const uint32_t idesc = make_idesc(FMT_TF32, 128, nr, false, S.b_mn != 0);
          for (int kb = 0; kb < nkb; ++kb) {
            uint32_t sa, sb;
            uint64_t* release;
            if (s == 0) {
              const int slot = kb % S0_RING;
              mbar_wait(&full0[slot], (kb / S0_RING) & 1);
              sa = smem_u32(pool + slot * S0_SLOT); sb = sa + 16384; release = &empty0[slot];
            } else {
              const int slot = it % B_RING;
              mbar_wait(&fullb[slot], (it / B_RING) & 1);
              sa = a_tile + kb * 16384; sb = smem_u32(bring + slot * B_SLOT); release = &emptyb[slot];
              ++it;
            }
            fence_after_thread_sync();
            // descriptors once per k-block; the four K = 8 steps only move the 14-bit start-address field (>> 4):
            // +32 B inside the swizzle row (K-major), +1024 B = 8 k-rows (MN-major).  One thread issues every MMA of the
            // chain, so its instruction count per MMA is chain latency.
            const uint64_t da0 = make_smem_desc(sa, 16, 1024, UMMA_SWIZZLE_128B);
            const uint64_t db0 = S.b_mn ? make_smem_desc(sb, 4096, 512, UMMA_SWIZZLE_128B_BASE32B)
                                        : make_smem_desc(sb, 16, 1024, UMMA_SWIZZLE_128B);
            const uint64_t db_step = S.b_mn ? 64u : 2u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_tf32(tmem_base + h * 128, da0 + (uint64_t)(2 * k), db0 + db_step * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            mma_commit(release);
          }
        }
        mma_commit(acc_full);
      }
    }
  } else {
    // ===== epilogue warps: thread = accumulator row; the two warps of a TMEM lane quarter alternate 32-column chunks
    // (the per-chunk work is a few hundred dependent ALU instructions: with one warp per scheduler it ran at ~1 us per
    // chunk, latency-bound)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int64_t grow = (int64_t)m0 + row;
    const bool valid = grow < a.bs;
    const bool issuer = threadIdx.x == 64;    // issues the TMA stores of the stage outputs
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int y_label = (valid && half == 0) ? a.labels[grow] : -1;      // fetched now, long before the softmax stage needs it
    // sign-bit word w of mask slot m for this row (slot 0: 8 words, slots 1..3: 4 words each)
    auto mask_word = [&](int m, int w) -> uint32_t* { return s_mask + ((m == 0 ? w : 8 + (m - 1) * 4 + w) * 128 + row); };
    // Sign bits of this row of the layer-1 activations (the mask of the LAST stage): picked up from the k-blocks as they
    // pass through the stage-0 ring -- nothing is re-read from global memory.  A slot is recycled only when the MMAs AND
    // these 128 readers are done with it (empty0 counts 129).
    if (half == 0) {
#pragma unroll 1
      for (int kb = 0; kb < nkb0; ++kb) {
        const int slot = kb % S0_RING;
        mbar_wait(&full0[slot], (kb / S0_RING) & 1);
        const uint8_t* arow = pool + slot * S0_SLOT + row * 128;
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 x = *reinterpret_cast<const float4*>(arow + ((j ^ (row & 7)) << 4));
          bits |= (x.x > 0.f ? 1u : 0u) << (4 * j) | (x.y > 0.f ? 1u : 0u) << (4 * j + 1) |
                  (x.z > 0.f ? 1u : 0u) << (4 * j + 2) | (x.w > 0.f ? 1u : 0u) << (4 * j + 3);
        }
        *mask_word(0, kb & 7) = bits;
        mbar_arrive(&empty0[slot]);
      }
    }
    if (blockIdx.x == 0 && issuer) step_bookkeeping(a.bk);
#pragma unroll 1
    for (int s = 0; s < a.nstages; ++s) {
      const Stage& S = a.st[s];
      const bool last = (s == a.nstages - 1);
      // the output tile: the activation buffer the next stage reads; the last stage's (up to 256 wide) spans both
      uint8_t* tile = last ? pool : pool + (s & 1) * ACT_BYTES;
      mbar_wait(acc_full, s & 1);
      fence_after_thread_sync();
      if (issuer) {
        trace_mark(a.trace, s == 0 ? 1 : (S.epi == 1 ? 2 : 3));
        // the TMA store that last read the buffer we are about to overwrite (two stages ago; for the last stage also
        // the previous one) must have finished reading it
        if (last) bulk_wait_read<0>(); else bulk_wait_read<1>();
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int nchunks = (S.epi == 1) ? 1 : ((S.N + 31) >> 5);
      if (S.epi == 1) {
        if (half == 0) {
          // ---- output layer: logits -> softmax -> CE, arg-max, dZ = p - y (one 32-column chunk)
          float v[32];
          tmem_ld32(t_lane, v);
          const int C = S.N;
          float mx = -INFINITY;
          int arg = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < C) {
              v[j] += s_bias[s * MAX_MID + j];
              if (v[j] > mx) { mx = v[j]; arg = j; }     // first maximum: np.argmax tie-break
            }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < C) { v[j] = expf(v[j] - mx); sum += v[j]; }
          const float inv = 1.0f / sum;
          const int y = y_label;
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
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tile + act_off(row, 4 * j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          double loss = valid ? -log((double)py + 1e-9) : 0.0;
          float correct = (valid && arg == y) ? 1.f : 0.f;
          loss = warp_sum_d(loss);
          correct = warp_sum(correct);
          if (lane == 0) stats_add(a.stats, loss, (unsigned long long)correct);
        }
      } else {
#pragma unroll 1
        for (int ch = half; ch < nchunks; ch += 2) {
          const int c = ch * 32;
          float v[32];
          tmem_ld32(t_lane + c, v);
          // columns of this chunk that exist (rows past the batch: none)
          const uint32_t colmask = !valid ? 0u : (S.N - c >= 32 ? 0xffffffffu : ((1u << (S.N - c)) - 1u));
          if (S.epi == 0) {
            uint32_t bits = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float x = v[j] + s_bias[s * MAX_MID + c + j];
              x = (x < 0.f) ? 0.f : x;                   // the reference's comparison: NaN propagates (forward_pass.pyx:8-13)
              x = ((colmask >> j) & 1u) ? x : 0.f;
              v[j] = x;
              bits |= (x > 0.f ? 1u : 0u) << j;
            }
            *mask_word(S.mask_slot & 3, ch & 3) = bits;
          } else {
            const uint32_t bits = *mask_word(S.mask_slot & 3, ch & 7) & colmask;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.f;     // z <= 0 -> 0 (backward_pass.pyx:86)
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tile + act_off(row, c + 4 * j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      // the tile is complete (generic-proxy writes -> async proxy) and TMEM is drained: the next stage's MMAs may go
      fence_proxy_async_smem();
      fence_before_thread_sync();
      mbar_arrive(act_ready);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // the same tile also goes to global memory, once, for the dW GEMMs and the bias-gradient column sums: one TMA store
      // per k-block (rows past the batch and columns past the layer width are clipped by the tensor map)
      if (issuer) {
        for (int ch = 0; ch < nchunks; ++ch) tma_store_2d(&maps.o[s], tile + ch * 16384, ch * 32, m0);
        bulk_commit();
      }
      if (S.epi != 0) {
        // Bias gradient of this row tile (backward_pass.pyx:72,93): thread t adds column t of the finished tile over its
        // 128 rows, top to bottom (a warp reads 32 consecutive words of one row: conflict-free), and stores ONE ordered
        // partial; rows past the batch hold zeros.  Runs beside the next stage's MMAs.
        const int t = (int)threadIdx.x - 64;
        if (t < (int)S.ld_colsum) {
          float acc = 0.f;
          if (t < S.N) {
            const uint8_t* col = tile + (t >> 5) * 16384 + ((t & 3) << 2);
            const int cidx = (t >> 2) & 7;
#pragma unroll 8
            for (int r = 0; r < 128; ++r) acc += *reinterpret_cast<const float*>(col + r * 128 + ((cidx ^ (r & 7)) << 4));
          }
          S.colsum_part[(int64_t)blockIdx.x * S.ld_colsum + t] = acc;
        }
      }
    }
    if (issuer) bulk_wait_read<0>();          // the stores have read their tiles: the shared memory may go
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
