// gemm_tc.cuh -- the dense-MLP GEMMs on 5th-gen tensor cores (K2 of SURVEY 2.3).
//
//   C[M,N] (+)= A . B     TF32 operands, FP32 accumulation in TMEM, tcgen05.mma cta_group::1
//
// Three operand-layout forms cover forward, dA and dW with row-major storage everywhere
// (forward_pass.pyx:29,36 / backward_pass.pyx:71,75,92,97):
//   forward  Z  = X . W^T   A = X  [M=bs , K=in ] K-major     B = W  [N=out, K=in ] K-major
//   dA       dA = dZ . W    A = dZ [M=bs , K=out] K-major     B = W  [K=out, N=in ] MN-major
//   dW       dW = dZ^T . X  A = dZ [K=bs , M=out] MN-major    B = X  [K=bs , N=in ] MN-major
//
// Pipeline (one 128 x BN output tile per CTA, optional split-K over gridDim.z):
//   warp 0 (1 thread)  TMA producer: cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier expect_tx
//   warp 1 (1 thread)  MMA issuer: 4 x tcgen05.mma (K=8 each) per 32-wide k-block, tcgen05.commit
//                      releases the smem stage / signals the epilogue
//   warps 2-5          epilogue: tcgen05.ld (32 lanes x 32 cols) -> bias / ReLU / ReLU-mask ->
//                      128-bit global stores; split-K slices that cannot meet in distributed shared memory store
//                      their raw partial tiles side by side (summed in slice order by a later pass: no atomics)
// K-major tiles are [rows][32 floats] with SWIZZLE_128B (SBO 1024 B); MN-major tiles are
// [MN/32 atoms][32 k-rows][32 floats] in the SWIZZLE_128B_BASE32B form -- the only smem layout the
// tensor core accepts for MN-major TF32 (32-B swizzle granules, 4 k-rows per atom; LBO = 4096 B between
// MN atoms, SBO = 512 B between 4-row groups) -- loaded by one {32 x 32} TMA box per atom from a plain
// 2-D map {MN, K} encoded with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
#pragma once
#include "elementwise.cuh"
#include "tc_common.cuh"

namespace pb {
namespace tc {

constexpr int G_BM = 128;
constexpr int G_BK = 32;  // floats: one 128-byte swizzle row

// TWO: the CTA pair of a cluster runs as one tensor-core unit (cta_group::2): a 256 x BN tile per pair, every CTA
// stages its own 128 rows of A and only HALF of the B tile -- 32 KB instead of 48 KB per k-block for BN = 256.  The
// TF32 GEMMs are bound by what an SM can take in per second (~93 GB/s measured: 48 KB per 128x256x32 step is twice
// what the tensor pipe needs time for), so the smaller stage is worth more than any multicast, which only relieves L2.
template <int BN, bool TWO = false>
struct GemmCfg {
  static constexpr int A_BYTES = G_BM * G_BK * 4;       // 16 KB
  static constexpr int B_BYTES = (TWO ? BN / 2 : BN) * G_BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // <= 208 KB of stages: clustered launches of a CTA that asks for more than ~200 KB of dynamic shared memory are
  // refused on this part ("cluster misconfiguration", seen at 225 KB), and 6 x 32 KB stages are deep enough
  static constexpr int MAX_STAGES = (208 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
  static constexpr int GATHER_IDX = 2048;                // row indices a CTA of the gather-free form stages (TWO only)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + (TWO ? GATHER_IDX * 4 : 0);
};

struct TcGemmArgs {
  float* C;
  int64_t ldc;
  int M, N, K;
  int epi;      // 0 none, 1 +bias, 2 relu(+bias), 3 mask (only when !partial)
  int partial;  // 1: split-K slice z stores its raw partial tile at C + z * part_stride (no epilogue); the slices are
                //    added in z order by grad_finalize_kernel (dW) or splitk_reduce_kernel (which applies the epilogue)
  int64_t part_stride;
  int num_kb;        // total 32-wide k-blocks
  int kb_per_split;  // k-blocks per gridDim.z slice
  const float* bias;
  const float* mask;
  int64_t ldmask;
  float* colsum_part;  // optional (bias gradient, !partial only): colsum_part[blockIdx.y * ld_colsum + n] = sum of the epilogue
  int64_t ld_colsum;   //   output over this CTA's 128 rows -- one ordered partial per row tile, added by grad_finalize_kernel
  int cluster_reduce;  // 1: the gridDim.z k-slices of a tile share a cluster; partial tiles meet in distributed shared memory
  int pair;            // 1: two M-adjacent CTAs share a cluster; each loads half of the B tile and multicasts it to both
                       //    (TWO kernels: always a pair -- set to 1 so that the shared cluster plumbing applies)
  // epi == 4 (output layer of a training step, N <= 32, one N tile): logits + bias -> softmax -> cross-entropy,
  // dZ = p - onehot, db_o, loss / #correct tallies and the per-step bookkeeping, all on the accumulator row a thread
  // already holds after tcgen05.ld (replaces softmax_ce_kernel; forward_pass.pyx:36-37, pycnn.py:634-640,
  // backward_pass.pyx:60,72)
  const int32_t* labels;
  float* dz;
  int64_t lddz;
  unsigned long long* stats;
  float* db_part;      // db_o partial of this row tile: db_part[blockIdx.y * lddz + j]
  StepBookkeeping bk;
  unsigned long long* trace;   // timeline slot (diagnostics) or nullptr
  // Gather-free layer-1 GEMMs (2-SM kernels only): the batch rows are rows idx[*cursor + extra + i] of the DATASET, fetched by
  // TMA tile::gather4 straight into the operand stage -- forward: A rows (map_a = dataset, box {32, 1}); dW: B k-rows
  // (map_b = dataset, box {32, 1}, 32-byte-atom swizzle).  No gathered copy of the batch ever exists in HBM.
  const int64_t* gather_idx;
  const int64_t* gather_cursor;
  int64_t gather_extra;
};

// ---- thread-block cluster helpers (split-K reduction through distributed shared memory) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_x() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctaid_y() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.y;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctaid_z() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctaid.z;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_y() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctaid.y;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctaid_z() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctaid.z;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_dsmem_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Cluster split-K tail (see the kernel): CTA `rank` of S finishes rows [rank*128/S, (rank+1)*128/S) of the tile.  A warp
// takes RG rows at a time so that RG*S = 16 remote 128-bit loads are in flight per lane (DSMEM latency is ~2x L2's).
template <int BN, int S>
__device__ __forceinline__ void cluster_reduce_rows(const TcGemmArgs& g, uint8_t* smem, int q, int lane, int m0, int n0,
                                                    uint32_t P, uint32_t py) {
  constexpr int PITCH = BN + 4;
  constexpr int RPC = G_BM / S;        // rows per CTA: 64 / 32 / 16
  constexpr int RPW = RPC / 4;         // rows per warp: 16 / 8 / 4
  constexpr int RG = 16 / S;           // rows per batch: 8 / 4 / 2
  constexpr int PASSES = (BN + 127) / 128;
  // cluster shape (1, P, S) -- or (P, 1, S) for the 2-SM kernels -- with P = 2 when M-adjacent CTAs pair up:
  // %cluster_ctarank = (pair index py) + P*z
  const uint32_t rank = cluster_ctaid_z();
  const int ncols = min(BN, g.N - n0);
  const uint32_t stage_local = smem_u32(smem);
  uint32_t stage_of[S];
#pragma unroll
  for (int z = 0; z < S; ++z) stage_of[z] = map_to_cta(stage_local, py + P * (uint32_t)z);
  const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                      (g.epi != 3 || (((g.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0)));
#pragma unroll 1
  for (int pass = 0; pass < PASSES; ++pass) {
    const int cc = pass * 128 + lane * 4;
    if (cc >= ncols) continue;
    const int n = n0 + cc;
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.epi == 1 || g.epi == 2) {
#pragma unroll
      for (int e = 0; e < 4; ++e) bias4[e] = (n + e < g.N) ? __ldg(g.bias + n + e) : 0.f;
    }
#pragma unroll 1
    for (int b = 0; b < RPW; b += RG) {
      // tile rows of this batch: rank*RPC + q*RPW + b + i   (a warp owns RPW consecutive rows)
      const int r0 = (int)rank * RPC + q * RPW + b;
      if ((int64_t)m0 + r0 >= g.M) break;
      float4 part[RG][S];
#pragma unroll
      for (int i = 0; i < RG; ++i) {
        const uint32_t off = (uint32_t)((r0 + i) * PITCH + cc) * 4u;
#pragma unroll
        for (int z = 0; z < S; ++z) part[i][z] = ld_dsmem_v4(stage_of[z] + off);
      }
      float mk[RG][4];
      if (g.epi == 3) {
#pragma unroll
        for (int i = 0; i < RG; ++i) {
          const int64_t grow = (int64_t)m0 + r0 + i;
          if (grow < g.M) {
            const float* mrow = g.mask + grow * g.ldmask + n;
#pragma unroll
            for (int e = 0; e < 4; ++e) mk[i][e] = (vec_ok || n + e < g.N) ? __ldg(mrow + e) : 0.f;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < RG; ++i) {
        const int64_t grow = (int64_t)m0 + r0 + i;
        if (grow < g.M) {
          float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int z = 0; z < S; ++z) { o[0] += part[i][z].x; o[1] += part[i][z].y; o[2] += part[i][z].z; o[3] += part[i][z].w; }
          if (g.epi == 1 || g.epi == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] += bias4[e];
          }
          if (g.epi == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (o[e] < 0.f) ? 0.f : o[e];
          }
          if (g.epi == 3) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (mk[i][e] <= 0.f) ? 0.f : o[e];
          }
          float* crow = g.C + grow * g.ldc + n;
          if (vec_ok) {
            *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (n + e < g.N) crow[e] = o[e];
          }
        }
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN, bool TWO = false>
__global__ void __launch_bounds__(192, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const TcGemmArgs g) {
  using Cfg = GemmCfg<BN, TWO>;
  constexpr int STAGES = Cfg::STAGES;
  // 128B-swizzled TMA/UMMA tiles need a 1024-byte aligned base; declaring the alignment (instead of fixing the
  // pointer up by hand) keeps every derived pointer in the shared address space (LDS/STS, not generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid: (N tiles, M tiles, k-slices); the 2-SM kernels put the M tiles on x, because the two CTAs of a cta_group::2
  // pair must be neighbours in the cluster's x dimension (a (1,2,1) cluster is refused at launch)
  const int tile_n = TWO ? blockIdx.y : blockIdx.x, tile_m = TWO ? blockIdx.x : blockIdx.y;
  const int n0 = tile_n * BN, m0 = tile_m * G_BM;
  const int kb0 = blockIdx.z * g.kb_per_split;
  const int kb1 = min(g.num_kb, kb0 + g.kb_per_split);
  const int nkb = kb1 - kb0;  // host guarantees >= 1

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], (g.pair && !TWO) ? 2 : 1);   // multicast pair: both CTAs' MMA issuers release a stage
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {                           // BN fp32 accumulator columns (power of two >= 32)
    if constexpr (TWO) tmem_alloc_2sm(tmem_slot, BN);
    else tmem_alloc(tmem_slot, BN);
  }
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  // paired CTAs write into each other's smem and arrive on each other's barriers: those must exist first
  uint16_t pair_mask = 0;
  uint32_t pair_half = 0;
  if (g.pair) {
    cluster_sync_all();
    const uint32_t r = cluster_ctarank();
    pair_mask = (uint16_t)(3u << (r & ~1u));
    pair_half = TWO ? cluster_ctaid_x() : cluster_ctaid_y();
  }
  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  pdl_wait();
  pdl_launch_dependents();
  trace_begin(g.trace);

  if (warp == 0) {
    bool gathered = false;
    if constexpr (TWO && A_MN == B_MN) {
      if (g.gather_idx) {
        // ===== TMA producer, gather-free form: the whole warp issues, one tile::gather4 (4 dataset rows x 32 floats) per
        // lane and k-block.  Forward: lane l owns batch rows m0 + 4l .. +3 of the A tile for the whole loop; dW: the B
        // tile of k-block i is batch rows 32 i .. +31, lane l = (N atom l / 8, row quad l % 8), indices staged in smem.
        gathered = true;
        const int64_t* idx = g.gather_idx + (g.gather_cursor ? g.gather_cursor[0] : 0) + g.gather_extra;
        int32_t* s_idx = reinterpret_cast<int32_t*>(smem + STAGES * Cfg::STAGE_BYTES + 256);
        int r4[4] = {-1, -1, -1, -1};
        if constexpr (!A_MN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row = m0 + 4 * lane + j;
            if (row < g.M) r4[j] = (int)idx[row];
          }
        } else {
          for (int i = lane; i < nkb * 32; i += 32) {
            const int row = kb0 * 32 + i;
            s_idx[i] = row < g.K ? (int)idx[row] : -1;
          }
          __syncwarp();
        }
        for (int i = 0; i < nkb; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (i / STAGES) & 1;
          if (lane == 0) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (pair_half == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          }
          __syncwarp();
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const int k0 = (kb0 + i) * G_BK;
          const uint32_t lead_full = map_to_cta(smem_u32(&full_bar[s]), cluster_ctarank() & ~1u);
          if constexpr (!A_MN) {
            tma_gather4_2sm(sa + lane * 512, &map_a, lead_full, k0, r4[0], r4[1], r4[2], r4[3]);
            if (lane == 0) tma_load_2d_2sm(sb, &map_b, lead_full, k0, n0 + (int)pair_half * (BN / 2));
          } else {
            if (lane < G_BM / 32) tma_load_2d_2sm(sa + lane * (G_BK * 128), &map_a, lead_full, m0 + lane * 32, k0);
            const int a2 = lane >> 3, gq = lane & 7;
            if (a2 < BN / 64) {
              const int32_t* ri = s_idx + i * 32 + gq * 4;
              tma_gather4_2sm(sb + a2 * (G_BK * 128) + gq * 512, &map_b, lead_full,
                              n0 + ((int)pair_half * (BN / 64) + a2) * 32, ri[0], ri[1], ri[2], ri[3]);
            }
          }
        }
      }
    }
    if (!gathered && lane == 0) {
      // ===== TMA producer =====
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        const int k0 = (kb0 + i) * G_BK;
        if constexpr (TWO) {
          // both CTAs of the pair fill their own stage (own A rows, own half of B) and report the bytes to the LEADER's
          // full barrier, which the leader armed for the pair's 2 x STAGE_BYTES
          const uint32_t lead_full = map_to_cta(smem_u32(&full_bar[s]), cluster_ctarank() & ~1u);
          if (pair_half == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int at = 0; at < G_BM / 32; ++at) tma_load_2d_2sm(sa + at * (G_BK * 128), &map_a, lead_full, m0 + at * 32, k0);
          } else {
            tma_load_2d_2sm(sa, &map_a, lead_full, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int a2 = 0; a2 < BN / 64; ++a2)
              tma_load_2d_2sm(sb + a2 * (G_BK * 128), &map_b, lead_full, n0 + ((int)pair_half * (BN / 64) + a2) * 32, k0);
          } else {
            tma_load_2d_2sm(sb, &map_b, lead_full, k0, n0 + (int)pair_half * (BN / 2));
          }
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        if (A_MN) {
#pragma unroll
          for (int at = 0; at < G_BM / 32; ++at) tma_load_2d(sa + at * (G_BK * 128), &map_a, &full_bar[s], m0 + at * 32, k0);
        } else {
          tma_load_2d(sa, &map_a, &full_bar[s], k0, m0);
        }
        if (g.pair) {
          // this CTA fetches its half of the B tile once and multicasts it into both CTAs of the pair (each
          // full barrier still sees STAGE_BYTES: own A + two B halves)
          if constexpr (BN >= 64) {
            if (B_MN) {
#pragma unroll
              for (int a2 = 0; a2 < BN / 64; ++a2) {
                const int at = (int)pair_half * (BN / 64) + a2;
                tma_load_2d_mc(sb + at * (G_BK * 128), &map_b, &full_bar[s], n0 + at * 32, k0, pair_mask);
              }
            } else {
              tma_load_2d_mc(sb + pair_half * (Cfg::B_BYTES / 2), &map_b, &full_bar[s], k0, n0 + (int)pair_half * (BN / 2), pair_mask);
            }
          }
        } else if (B_MN) {
#pragma unroll
          for (int at = 0; at < BN / 32; ++at) tma_load_2d(sb + at * (G_BK * 128), &map_b, &full_bar[s], n0 + at * 32, k0);
        } else {
          tma_load_2d(sb, &map_b, &full_bar[s], k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (!TWO || pair_half == 0)) {
      // ===== MMA issuer (TWO: the leader CTA issues for the pair, M = 256) =====
      constexpr uint32_t idesc = make_idesc(FMT_TF32, TWO ? 2 * G_BM : G_BM, BN, A_MN, B_MN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        fence_after_thread_sync();
        if (i == 0) trace_mark(g.trace, 1);                       // first operand stage landed
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < G_BK / 8; ++k) {
          // K-major: advance 8 floats (32 B) inside the 128-B swizzle row; MN-major: next 8 k-rows (1024 B)
          // MN-major TF32 operands only exist in the SWIZZLE_128B_BASE32B form (32-byte swizzle granules,
          // 4 k-rows per 512-B atom): LBO = stride between 32-float MN atoms, SBO = stride between 4-row groups
          constexpr uint32_t mn_lbo = G_BK * 128, mn_sbo = 512u;
          const uint64_t da = A_MN ? make_smem_desc(sa + k * 1024, mn_lbo, mn_sbo, UMMA_SWIZZLE_128B_BASE32B)
                                   : make_smem_desc(sa + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
          const uint64_t db = B_MN ? make_smem_desc(sb + k * 1024, mn_lbo, mn_sbo, UMMA_SWIZZLE_128B_BASE32B)
                                   : make_smem_desc(sb + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
          if constexpr (TWO) mma_tf32_2sm(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          else mma_tf32(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        // frees this smem stage once the MMAs above have read it (in both CTAs of a pair: the peer's next
        // multicast writes into this CTA's stage too / the peer's stage was read by this pair MMA)
        if constexpr (TWO) mma_commit_2sm_mc(&empty_bar[s], pair_mask);
        else if (g.pair) mma_commit_mc(&empty_bar[s], pair_mask);
        else             mma_commit(&empty_bar[s]);
      }
      if constexpr (TWO) mma_commit_2sm_mc(tmem_full_bar, pair_mask);   // accumulator complete, in both CTAs
      else mma_commit(tmem_full_bar);
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
    // 1. tcgen05.ld gives thread i the 32-column chunks of accumulator row i; they are parked in this warp's
    //    staging tile in the (now idle) operand smem with a pitch of BN+4 floats (conflict-free 128-bit stores);
    // 2. the warp then walks its 32 rows with lanes along the columns, so bias / mask loads, stores and split-K
    //    reductions are all coalesced 128-bit accesses; column sums for the bias gradient are taken on the way.
    const int q = warp & 3;
    constexpr int PITCH = BN + 4;
    float* stage = reinterpret_cast<float*>(smem) + q * 32 * PITCH;
    mbar_wait(tmem_full_bar, 0);
    fence_after_thread_sync();
    if (threadIdx.x == 64) trace_mark(g.trace, 2);                // accumulator complete: the epilogue starts
    bool fused_softmax = false;
    if constexpr (BN == 32 && !A_MN && !B_MN) {
      if (g.epi == 4) {
        fused_softmax = true;
        if (tile_n == 0 && tile_m == 0 && q == 0 && lane == 0) step_bookkeeping(g.bk);   // per-step housekeeping
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16), v);
        const int C = g.N;
        const int64_t row = (int64_t)m0 + q * 32 + lane;
        const bool valid = row < g.M;
        float mx = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < C) {
            v[j] += __ldg(g.bias + j);
            if (v[j] > mx) { mx = v[j]; arg = j; }     // first maximum: np.argmax tie-break
          }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < C) { v[j] = expf(v[j] - mx); sum += v[j]; }
        const float inv = 1.0f / sum;
        const int y = valid ? g.labels[row] : -1;
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
          float4* drow = reinterpret_cast<float4*>(g.dz + row * g.lddz);   // lddz = pad4(C) <= 32, rows 16-byte aligned
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < (int)g.lddz) drow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        double loss = valid ? -log((double)py + 1e-9) : 0.0;
        float correct = (valid && arg == y) ? 1.f : 0.f;
        loss = warp_sum_d(loss);
        correct = warp_sum(correct);
        if (lane == 0) stats_add(g.stats, loss, (unsigned long long)correct);
        // db_o = column sums of dZ: per warp by shuffles, the four warps of the row tile meet in smem in warp order
        // and the CTA stores ONE ordered partial (grad_finalize_kernel adds the row tiles in order)
        float* cs_smem = reinterpret_cast<float*>(smem);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < C) {
            const float cs = warp_sum(v[j]);
            if (lane == 0) cs_smem[q * 32 + j] = cs;
          }
        asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps
        if (q == 0 && lane < (int)g.lddz) {
          const float tot = (lane < C) ? ((cs_smem[lane] + cs_smem[32 + lane]) + (cs_smem[64 + lane] + cs_smem[96 + lane])) : 0.f;
          g.db_part[(int64_t)tile_m * g.lddz + lane] = tot;
        }
      }
    }
    if (!fused_softmax) {
    const int ncols = min(BN, g.N - n0);             // valid columns of this tile (warp-uniform)
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      if (c >= ncols) break;
      float v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      float4* dst = reinterpret_cast<float4*>(stage + lane * PITCH + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncwarp();
    if (g.partial && !g.cluster_reduce) {
      // raw partial tile of a k-slice: every lane hands its staged row to the bulk-copy engine (one 1-D TMA store of
      // up to 1 KB per row) instead of walking it through registers -- the epilogue is then the TMEM drain alone
      float* const Cout = g.C + (int64_t)blockIdx.z * g.part_stride;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      fence_proxy_async_smem();
      __syncwarp();
      if (row < g.M) bulk_store(Cout + row * g.ldc + n0, stage + lane * PITCH, (uint32_t)(((ncols + 3) & ~3) * 4));
      bulk_commit();
      bulk_wait_read<0>();
    } else if (!g.cluster_reduce) {
    float* const Cout = g.C;
    const int row0 = m0 + q * 32;
    const int nrows = min(32, g.M - row0);            // warp-uniform, may be <= 0
    const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                        (g.epi != 3 || (((g.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.mask) & 15) == 0)));
    // lane -> (row, 4 columns): LPR lanes cover one row segment, RPI rows are handled per iteration and U
    // iterations are batched so that U independent mask loads are in flight per lane
    constexpr int LPR = (BN >= 128) ? 32 : BN / 4;
    constexpr int RPI = 32 / LPR;
    constexpr int U = (RPI == 1) ? 8 : 4;
    constexpr int PASSES = (BN + 127) / 128;
    const int lr = lane / LPR, lc = (lane % LPR) * 4;
#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
      const int cc = pass * 128 + lc;
      const bool col_ok = cc < ncols;
      const int n = n0 + cc;
      float bias4[4] = {0.f, 0.f, 0.f, 0.f};
      if (col_ok && !g.partial && (g.epi == 1 || g.epi == 2)) {
#pragma unroll
        for (int e = 0; e < 4; ++e) bias4[e] = (n + e < g.N) ? __ldg(g.bias + n + e) : 0.f;
      }
      float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int rb = 0; rb < 32; rb += U * RPI) {
        float mk[U][4];
        if (g.epi == 3 && !g.partial) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int r = rb + u * RPI + lr;
            if (col_ok && r < nrows) {
              const float* mrow = g.mask + (int64_t)(row0 + r) * g.ldmask + n;
              if (vec_ok) {
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(mrow));
                mk[u][0] = m4.x; mk[u][1] = m4.y; mk[u][2] = m4.z; mk[u][3] = m4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) mk[u][e] = (n + e < g.N) ? __ldg(mrow + e) : 0.f;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = rb + u * RPI + lr;
          if (col_ok && r < nrows) {
            const float4 acc = *reinterpret_cast<const float4*>(stage + r * PITCH + cc);
            float o[4] = {acc.x, acc.y, acc.z, acc.w};
            float* crow = Cout + (int64_t)(row0 + r) * g.ldc + n;
            if (!g.partial) {
              if (g.epi == 1 || g.epi == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] += bias4[e];
              }
              if (g.epi == 2) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (o[e] < 0.f) ? 0.f : o[e];
              }
              if (g.epi == 3) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = (mk[u][e] <= 0.f) ? 0.f : o[e];
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) cs[e] += o[e];
            }
            if (vec_ok) {   // columns in [N, round_up(N,4)) are padding of the row and receive exact zeros
              *reinterpret_cast<float4*>(crow) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (n + e < g.N) crow[e] = o[e];
            }
          }
        }
      }
      if (g.colsum_part && !g.partial) {   // db[n] = sum over this warp's 32 rows (backward_pass.pyx:93)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], o);
        }
        // per-warp sums meet in smem (behind the staging tiles) in warp order; the CTA stores one ordered partial per
        // column (same-address atomics of 32 M-tiles x 4 warps were also a 2-4 us completion tail of the dA GEMMs)
        if (lr == 0 && col_ok) {
          float* cs_smem = reinterpret_cast<float*>(smem) + 128 * PITCH + q * BN + cc;
#pragma unroll
          for (int e = 0; e < 4; ++e) cs_smem[e] = cs[e];       // rows past M contributed nothing: zeros
        }
      }
    }
    if (g.colsum_part && !g.partial) {
      asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps
      const float* cs_all = reinterpret_cast<const float*>(smem) + 128 * PITCH;
      for (int j = (warp - 2) * 32 + lane; j < ncols; j += 128) {
        const float tot = (cs_all[j] + cs_all[BN + j]) + (cs_all[2 * BN + j] + cs_all[3 * BN + j]);
        if (n0 + j < g.N) g.colsum_part[(int64_t)tile_m * g.ld_colsum + n0 + j] = tot;
      }
    }
    }  // !cluster_reduce
    }  // !fused_softmax
  }
  if (g.cluster_reduce) {
    if (threadIdx.x == 64) trace_mark(g.trace, 3);                // partial tile parked in smem
    // Split-K without global partial sums: the gridDim.z CTAs of one output tile form a thread-block cluster, each
    // has parked its raw 128 x BN partial tile in its own shared memory (above).  After a cluster barrier CTA r sums
    // rows [r*128/S, (r+1)*128/S) over the S partials in rank order (deterministic) straight out of its peers'
    // shared memory, applies bias / ReLU / mask and writes the finished rows of C -- no zero-fill, no atomics, no
    // fix-up pass.  Every thread of every CTA takes part in both barriers; the second keeps a CTA's shared memory
    // alive until its peers are done reading it.  (A push variant -- st.shared::cluster from TMEM into the owner's
    // smem -- measured 2.4 us slower: it needs one more cluster barrier before anyone may overwrite operand smem.)
    cluster_sync_all();
    if (warp >= 2) {
      const uint32_t P = TWO ? cluster_nctaid_x() : cluster_nctaid_y(), py = TWO ? cluster_ctaid_x() : cluster_ctaid_y();
      switch (cluster_nctaid_z()) {
        case 2: cluster_reduce_rows<BN, 2>(g, smem, warp & 3, lane, m0, n0, P, py); break;
        case 4: cluster_reduce_rows<BN, 4>(g, smem, warp & 3, lane, m0, n0, P, py); break;
        default: cluster_reduce_rows<BN, 8>(g, smem, warp & 3, lane, m0, n0, P, py); break;
      }
    }
    cluster_sync_all();
  } else if (g.pair) {
    cluster_sync_all();   // the peer's last stage releases arrive on this CTA's barriers: stay until it is done too
  }
  fence_before_thread_sync();
  __syncthreads();
  trace_end(g.trace);
  if (warp == 1) {
    fence_after_thread_sync();
    if constexpr (TWO) tmem_dealloc_2sm(tmem_base, BN);
    else tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int resolve_encode_tiled(pycnn_b200_ctx* c) {
  if (c->encode_tiled) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  PB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  PB_CHECK(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
  c->encode_tiled = fn;
  return 0;
}

// K-major operand: matrix [rows, K] row-major with pitch ld (floats).  Box {32 floats, box_rows}.
static inline int make_map_kmajor(pycnn_b200_ctx* c, CUtensorMap* map, const float* base, int64_t rows, int64_t K,
                                  int64_t ld, int box_rows) {
  PB_TRY(resolve_encode_tiled(c));
  PB_CHECK((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(c->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(K-major rows=%lld K=%lld ld=%lld) failed: %d", (long long)rows, (long long)K, (long long)ld, (int)r);
  return 0;
}

// MN-major operand: matrix [K, MN] row-major with pitch ld.  Plain 2-D map {MN (inner), K}; the kernel issues
// one {32 floats, 32 k-rows} box per 32-wide MN atom so smem holds [atom][k][32 floats] (128-B swizzled rows).
static inline int make_map_mnmajor(pycnn_b200_ctx* c, CUtensorMap* map, const float* base, int64_t K, int64_t MN,
                                   int64_t ld) {
  PB_TRY(resolve_encode_tiled(c));
  PB_CHECK((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)MN, (cuuint64_t)K};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(c->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(MN-major K=%lld MN=%lld ld=%lld) failed: %d", (long long)K, (long long)MN, (long long)ld, (int)r);
  return 0;
}

// Row-gather maps (tile::gather4): the dataset [rows, width] row-major with a box of ONE row x 32 floats; K-major operand
// tiles use the plain 128-byte swizzle, MN-major ones the 32-byte-atom form (same smem layouts as the tiled boxes).
static inline int make_map_gather(pycnn_b200_ctx* c, CUtensorMap* map, const float* base, int64_t rows, int64_t width, int64_t ld,
                                  bool mn_major) {
  PB_TRY(resolve_encode_tiled(c));
  PB_CHECK((ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA operand needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(c->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es,
      CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(row gather rows=%lld width=%lld) failed: %d", (long long)rows, (long long)width, (int)r);
  return 0;
}

struct GemmCall {
  const float* A; int64_t lda; bool a_kmajor;  // kmajor: [M,K] row-major; else [K,M] row-major
  const float* B; int64_t ldb; bool b_kmajor;  // kmajor: [N,K] row-major; else [K,N] row-major
  float* C; int64_t ldc;
  int M, N, K;
  int epi;
  const float* bias;
  const float* mask; int64_t ldmask;
  int cls;  // KernelClass
  // Gradient outputs are ordered partial sums (elementwise.cuh: GradParts), 1-based tensor index, 0 = none:
  int grad_slot;      // C is the gradient of tensor grad_slot-1: split-K slices go side by side into its arena region
  int colsum_slot;    // fused column sums of the output (bias gradient of tensor colsum_slot-1): one partial per row tile
  // epi == 4: fused softmax / cross-entropy / dZ epilogue of the output layer (C is not written); db_slot = bias tensor
  const int32_t* labels; float* dz; int64_t lddz; unsigned long long* stats; int db_slot; StepBookkeeping bk;
  int max_ctas;       // > 0: cap tiles x split (a GEMM meant to run on the SMs another kernel leaves idle)
  // gather-free form (layer-1 forward: A; layer-1 dW: B): the operand is the DATASET [gather_rows, K|N] and its batch rows
  // are rows gather_idx[*gather_cursor + gather_extra + i]; only honoured on the 2-SM path (gather_ok tells the caller)
  const int64_t* gather_idx; const int64_t* gather_cursor; int64_t gather_extra; int64_t gather_rows;
};

// how many clusters of S CTAs (S = 2, 4, 8 -> slot 0, 1, 2) of each instantiation the device can hold at once;
// queried once per process, outside any stream capture
static int g_max_clusters[4][3][2][4];   // [BN slot][layout slot][paired][S slot: 1, 2, 4, 8]
static inline int s_slot(int S) { return S == 1 ? 0 : S == 2 ? 1 : S == 4 ? 2 : 3; }
static inline int bn_slot(int BN) { return BN == 32 ? 0 : BN == 64 ? 1 : BN == 128 ? 2 : 3; }
static inline int layout_slot(bool a_mn, bool b_mn) { return !a_mn ? (b_mn ? 1 : 0) : 2; }

static int g_max_clusters2[3][3];   // cta_group::2 kernels (BN = 256): [layout slot][S slot: 1, 2, 4] clusters of (2, 1, S)

template <bool A_MN, bool B_MN>
static int prepare_tc_gemm_two() {
  auto kern = tc::gemm_tf32_kernel<256, A_MN, B_MN, true>;
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<256, true>::SMEM_BYTES));
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  for (int si = 0; si < 3; ++si) {
    const int S = 1 << si;
    int n = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(64, 1, S);
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = tc::GemmCfg<256, true>::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = S;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
    g_max_clusters2[layout_slot(A_MN, B_MN)][si] = n;
  }
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int prepare_tc_gemm() {
  auto kern = tc::gemm_tf32_kernel<BN, A_MN, B_MN>;
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::GemmCfg<BN>::SMEM_BYTES));
  // every kernel of the training step asks for the same L1/shared split: an SM only hosts CTAs of kernels whose
  // carve-out matches, so mixed preferences would serialise the PDL / fork-join / prefetch overlaps
  PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  for (int P = 1; P <= 2; ++P)
    for (int si = 0; si < 4; ++si) {
      const int S = 1 << si;
      int n = 0;
      if (P * S > 1 && P * S <= 8) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(1, 64, S);
        cfg.blockDim = dim3(192);
        cfg.dynamicSmemBytes = tc::GemmCfg<BN>::SMEM_BYTES;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = P; attr[0].val.clusterDim.z = S;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { n = 0; (void)cudaGetLastError(); }
      }
      g_max_clusters[bn_slot(BN)][layout_slot(A_MN, B_MN)][P - 1][si] = n;
    }
  return 0;
}
template <bool A_MN, bool B_MN>
static int prepare_tc_gemm_bn() {
  PB_TRY((prepare_tc_gemm<32, A_MN, B_MN>()));
  PB_TRY((prepare_tc_gemm<64, A_MN, B_MN>()));
  PB_TRY((prepare_tc_gemm<128, A_MN, B_MN>()));
  PB_TRY((prepare_tc_gemm<256, A_MN, B_MN>()));
  return 0;
}
// opt every instantiation into > 48 KB of dynamic smem once per device, outside any stream capture
static int tc_gemm_prepare(pycnn_b200_ctx*) {
  PB_TRY((prepare_tc_gemm_bn<false, false>()));
  PB_TRY((prepare_tc_gemm_bn<false, true>()));
  PB_TRY((prepare_tc_gemm_bn<true, true>()));
  PB_TRY((prepare_tc_gemm_two<false, false>()));
  PB_TRY((prepare_tc_gemm_two<false, true>()));
  PB_TRY((prepare_tc_gemm_two<true, true>()));
  return 0;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_tc_gemm(pycnn_b200_ctx* c, const GemmCall& g, const CUtensorMap& ma, const CUtensorMap& mb,
                          const tc::TcGemmArgs& args, dim3 grid) {
  using Cfg = tc::GemmCfg<BN>;
  auto kern = tc::gemm_tf32_kernel<BN, A_MN, B_MN>;
  const int cy = args.pair ? 2 : 1, cz = args.cluster_reduce ? (int)grid.z : 1;
  PB_CUDA(launch_pdl_cluster(c, kern, grid, dim3(192), (size_t)Cfg::SMEM_BYTES, c->stream, cy, cz, ma, mb, args));
  return 0;
}

template <bool A_MN, bool B_MN>
static int dispatch_bn(pycnn_b200_ctx* c, int BN, const GemmCall& g, const CUtensorMap& ma, const CUtensorMap& mb,
                       const tc::TcGemmArgs& args, dim3 grid, bool two) {
  if (two) {
    auto kern = tc::gemm_tf32_kernel<256, A_MN, B_MN, true>;
    // M tiles on x: the CTAs of a cta_group::2 pair must be neighbours in the cluster's x dimension
    PB_CUDA(launch_pdl_cluster3(c, kern, dim3(grid.y, grid.x, grid.z), dim3(192), (size_t)tc::GemmCfg<256, true>::SMEM_BYTES,
                                c->stream, 2, 1, args.cluster_reduce ? (int)grid.z : 1, ma, mb, args));
    return 0;
  }
  switch (BN) {
    case 32: return launch_tc_gemm<32, A_MN, B_MN>(c, g, ma, mb, args, grid);
    case 64: return launch_tc_gemm<64, A_MN, B_MN>(c, g, ma, mb, args, grid);
    case 128: return launch_tc_gemm<128, A_MN, B_MN>(c, g, ma, mb, args, grid);
    default: return launch_tc_gemm<256, A_MN, B_MN>(c, g, ma, mb, args, grid);
  }
}

struct GemmPlan {
  int BN, tiles_m, tiles_n, num_kb, split, kb_per_split;
  bool cluster, pair, two;
};

// Tiling of one GEMM: a pure function of the shape, the call's epilogue and the context's switches (ensure_workspace
// sizes the partial-sum arena from the same function).
static GemmPlan plan_tc_gemm(const pycnn_b200_ctx* c, const GemmCall& g) {
  GemmPlan p{};
  int BN = g.N <= 32 ? 32 : g.N <= 64 ? 64 : g.N <= 128 ? 128 : 256;
  const int tiles_m = (int)ceil_div(g.M, tc::G_BM);
  const int num_kb = (int)ceil_div(g.K, tc::G_BK);
  // shallow-K problems are epilogue/latency bound: prefer narrower tiles so that more SMs share the epilogue
  // (the A tile is then re-read from L2 once per N tile, which is cheap when K is small)
  // -- but not past one wave: a second wave of 128 x 32 tiles costs more than twice-as-wide tiles in a single one
  if (num_kb <= 16) {
    while (BN > 32 && tiles_m * (int)ceil_div(g.N, BN) < c->num_sms) BN >>= 1;
    if (c->one_wave_tiles && BN < 256 && tiles_m * (int)ceil_div(g.N, BN) > c->num_sms &&
        tiles_m * (int)ceil_div(g.N, BN * 2) * 2 > c->num_sms)
      BN <<= 1;
  }
  if (c->force_bn > 0 && num_kb > 16 && g.N >= c->force_bn) BN = c->force_bn;   // experiments (PYCNN_B200_FORCE_BN)
  const int tiles_n = (int)ceil_div(g.N, BN);
  // deep-K problems whose tile grid cannot fill the 148 SMs: split K across gridDim.z
  int split = 1;
  const int tiles = tiles_m * tiles_n;
  if (tiles * 2 <= c->num_sms && num_kb >= 16 && !g.colsum_slot && g.epi != 4) {
    split = (int)std::min<int64_t>(c->num_sms / tiles, num_kb / 4);
    if (split < 1) split = 1;
  }
  if (g.max_ctas > 0) split = std::max(1, std::min(split, g.max_ctas / std::max(1, tiles)));
  int kb_per_split = (int)ceil_div(num_kb, split);
  split = (int)ceil_div(num_kb, kb_per_split);  // no empty slices
  // When the output needs an epilogue the k-slices of a tile run as one thread-block cluster and meet in distributed
  // shared memory: no workspace, no second pass.  Needs a cluster size of 2/4/8 with no empty slice and all clusters
  // resident at once.  Gradient outputs (grad_slot) keep the wide split and store ordered partials instead.
  // Deep-K problems with at least two M tiles also pair M-adjacent CTAs: both need the same B tile, each fetches
  // half of it and multicasts it to the pair (a third less L2->SM traffic per k-block for a 128 x 256 tile).
  const auto& occ = g_max_clusters[bn_slot(BN)][layout_slot(!g.a_kmajor, !g.b_kmajor)];
  // (measured: +30% on 500-k-block loops, nothing on 32-k-block ones, where the cluster launch costs as much as it saves)
  const bool pair_ok = c->use_pair && tiles_m >= 2 && BN >= 64 && kb_per_split >= c->pair_min_kb && g.epi != 4;
  const int pair_tiles = tiles_n * (int)ceil_div(tiles_m, 2);
  bool cluster = false, pair = false, two = false;
  // Deep-K 256-wide tiles with at least two M tiles run as CTA pairs on the 2-SM tensor-core path (cta_group::2).
  // Split-K slices of such pairs meet in distributed shared memory when the (1, 2, S) clusters all fit, else they
  // store ordered partials (gradients: the arena; anything else: the shared workspace + splitk_reduce_kernel).
  const auto& occ2 = g_max_clusters2[layout_slot(!g.a_kmajor, !g.b_kmajor)];
  if (c->use_two_sm && BN == 256 && tiles_m >= 2 && kb_per_split >= 8 && g.epi != 4 && !g.colsum_slot &&
      occ2[0] >= std::min(pair_tiles * split, c->num_sms / 2)) {
    two = true;
    if (split > 1 && !g.grad_slot && c->use_cluster_splitk && (split == 2 || split == 4) && occ2[s_slot(split)] >= pair_tiles)
      cluster = true;
  } else if (split > 1 && !g.grad_slot && c->use_cluster_splitk) {
    for (int si = 3; si >= 1 && !cluster; --si) {
      const int S = 1 << si;
      if (S > split) continue;
      const int kbps = (int)ceil_div(num_kb, S);
      if ((int)ceil_div(num_kb, kbps) != S) continue;
      if (pair_ok && 2 * S <= 8 && occ[1][si] >= pair_tiles) { cluster = pair = true; }
      else if (occ[0][si] >= tiles) { cluster = true; }
      else continue;
      split = S; kb_per_split = kbps;
    }
  }
  if (!two && !cluster && pair_ok && occ[1][0] >= std::min(pair_tiles * split, c->num_sms / 2)) pair = true;
  p.two = two;
  p.BN = BN; p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.num_kb = num_kb; p.split = split; p.kb_per_split = kb_per_split;
  p.cluster = cluster; p.pair = pair;
  return p;
}

// can this call run gather-free (2-SM path, indices of a k-slice fit the kernel's staging)?
static bool gather_ok(const pycnn_b200_ctx* c, const GemmCall& g) {
  const GemmPlan p = plan_tc_gemm(c, g);
  const bool dw = !g.a_kmajor && !g.b_kmajor;
  return p.two && (g.a_kmajor == g.b_kmajor) && (!dw || p.kb_per_split * 32 <= tc::GemmCfg<256, true>::GATHER_IDX);
}

static int tc_gemm(pycnn_b200_ctx* c, const GemmCall& g) {
  PB_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "tc_gemm: empty problem");
  PB_CHECK(!( !g.a_kmajor && g.b_kmajor), "tc_gemm: (A MN-major, B K-major) is not instantiated");
  PB_CHECK(g.epi != 4 || (g.N <= 32 && g.a_kmajor && g.b_kmajor && g.labels && g.dz && g.stats && g.db_slot && (g.lddz & 3) == 0 && g.lddz <= 32),
           "tc_gemm: the fused softmax epilogue needs N <= 32 and K-major operands");
  const GemmPlan p = plan_tc_gemm(c, g);
  const int BN = p.BN, split = p.split;
  const bool partial = split > 1 && !p.cluster;
  CUtensorMap ma, mb;
  const bool gather_a = g.gather_idx && g.a_kmajor && g.b_kmajor, gather_b = g.gather_idx && !g.a_kmajor && !g.b_kmajor;
  PB_CHECK(!g.gather_idx || (p.two && (gather_a || gather_b) && (gather_a || p.kb_per_split * 32 <= tc::GemmCfg<256, true>::GATHER_IDX)),
           "tc_gemm: the gather-free form needs the 2-SM path (ask gather_ok first)");
  if (gather_a)        PB_TRY(make_map_gather(c, &ma, g.A, g.gather_rows, g.K, g.lda, false));
  else if (g.a_kmajor) PB_TRY(make_map_kmajor(c, &ma, g.A, g.M, g.K, g.lda, tc::G_BM));
  else                 PB_TRY(make_map_mnmajor(c, &ma, g.A, g.K, g.M, g.lda));
  if (gather_b)        PB_TRY(make_map_gather(c, &mb, g.B, g.gather_rows, g.N, g.ldb, true));
  else if (g.b_kmajor) PB_TRY(make_map_kmajor(c, &mb, g.B, g.N, g.K, g.ldb, (p.pair || p.two) ? BN / 2 : BN));
  else                 PB_TRY(make_map_mnmajor(c, &mb, g.B, g.K, g.N, g.ldb));
  dim3 grid(p.tiles_n, (p.pair || p.two) ? (unsigned)round_up(p.tiles_m, 2) : p.tiles_m, split);
  tc::TcGemmArgs a{};
  a.C = g.C; a.ldc = g.ldc; a.M = g.M; a.N = g.N; a.K = g.K;
  a.epi = g.epi; a.partial = partial ? 1 : 0;
  a.num_kb = p.num_kb; a.kb_per_split = p.kb_per_split;
  a.bias = g.bias; a.mask = g.mask; a.ldmask = g.ldmask;
  a.cluster_reduce = p.cluster ? 1 : 0;
  a.pair = (p.pair || p.two) ? 1 : 0;
  a.labels = g.labels; a.dz = g.dz; a.lddz = g.lddz; a.stats = g.stats; a.bk = g.bk;
  a.gather_idx = g.gather_idx; a.gather_cursor = g.gather_cursor; a.gather_extra = g.gather_extra;
  const int64_t tile_floats = (int64_t)(g.M - 1) * g.ldc + round_up(g.N, 4);
  if (g.grad_slot) {                      // gradient tensor: ordered partials in its arena region (or C itself when unsplit)
    GradPartSlot& gs = c->gp[g.grad_slot - 1];
    gs.nparts = 0;
    if (partial) {
      PB_CHECK(split <= gs.cap && tile_floats <= gs.stride, "gradient partial arena too small for tensor %d (split %d, cap %d)", g.grad_slot - 1, split, gs.cap);
      gs.nparts = split;
      a.C = c->d_gpart + gs.off; a.part_stride = gs.stride;
    }
  } else if (partial) {                   // forward / dA fallback: shared workspace + ordered reduce with the epilogue below
    a.part_stride = round_up(tile_floats, 4);
    PB_CHECK(c->d_gemm_ws && (size_t)(a.part_stride * split) * sizeof(float) <= c->gemm_ws_bytes, "split-K workspace too small");
    PB_CHECK(c->stream != c->side_stream, "the shared split-K workspace belongs to the main stream");
    a.C = c->d_gemm_ws;
  }
  if (g.colsum_slot) {
    GradPartSlot& gs = c->gp[g.colsum_slot - 1];
    PB_CHECK((int)grid.y <= gs.cap, "bias-gradient partial arena too small for tensor %d (%d row tiles, cap %d)", g.colsum_slot - 1, (int)grid.y, gs.cap);
    gs.nparts = (int)grid.y;
    a.colsum_part = c->d_gpart + gs.off; a.ld_colsum = gs.stride;
  }
  if (g.epi == 4) {
    GradPartSlot& gs = c->gp[g.db_slot - 1];
    PB_CHECK((int)grid.y <= gs.cap && gs.stride == g.lddz, "output-bias partial arena does not fit (%d row tiles, cap %d)", (int)grid.y, gs.cap);
    gs.nparts = (int)grid.y;
    a.db_part = c->d_gpart + gs.off;
  }
  a.trace = trace_slot(c, g.cls);
  c->last_gemm_ctas = (int)(grid.x * grid.y * grid.z);
  {
    LaunchScope ls(c, g.cls);
    if (g.a_kmajor && g.b_kmajor) PB_TRY((dispatch_bn<false, false>(c, BN, g, ma, mb, a, grid, p.two)));
    else if (g.a_kmajor && !g.b_kmajor) PB_TRY((dispatch_bn<false, true>(c, BN, g, ma, mb, a, grid, p.two)));
    else PB_TRY((dispatch_bn<true, true>(c, BN, g, ma, mb, a, grid, p.two)));
  }
  if (partial && !g.grad_slot) {
    LaunchScope ls(c, KC_SPLITK);
    const int64_t total = (int64_t)g.M * ceil_div(g.N, 4);
    const int blocks = (int)std::min<int64_t>(ceil_div(total, 256), 8 * c->num_sms);
    PB_CUDA(launch_pdl(c, splitk_reduce_kernel, dim3(blocks), dim3(256), 0, c->stream, (const float*)c->d_gemm_ws, split,
                       a.part_stride, g.C, g.M, g.N, g.ldc, g.epi, g.bias, g.mask, g.ldmask));
  }
  return 0;
}

}  // namespace pb
