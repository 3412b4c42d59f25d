// extract_tc.cuh -- the fused feature extractor on tcgen05 tensor cores (K1 of SURVEY 2.3).
//
//   sum_c conv3x3_valid(flip(k_f), ch_c) -> ReLU -> 2x2 max-pool -> flatten      (pycnn/pycnn.py:273-287)
//
// Why tensor cores for a stencil: per 32x32 image the SIMT form needs 154 k FMAs (19 filters x 9 taps x
// 900 conv outputs) against 20 KB of HBM traffic, so FP32 issue -- not HBM -- bounds it.
// The contraction form removes that bound:
//   * a pooled output (py,px) depends on the 4x4 channel-summed patch P[u][v] = S[2py+u][2px+v];
//   * conv output of pool-quadrant q=(qy,qx) and filter f is  sum_{u,v} P[u][v] * Bq_f[u][v]  with
//     Bq_f[u][v] = flip(k_f)[u-qy][v-qx] (zero outside 0..2) -- a [16] x [16 x 4F] contraction;
//   * so ONE tcgen05.mma (M=128 pooled positions, N=4*F columns, K=16, kind::f16) produces the conv
//     outputs of all four quadrants and all filters for 128 pooled positions, accumulating in TMEM.
// Exactness: channel sums are integers <= 765 (exact in fp16); every tap is split hi+lo into two fp16
// values after a per-filter power-of-two normalisation (two MMAs into the same accumulator), products
// are exact in fp32 and the /255 of pycnn.py:273 is applied once, in fp32, in the epilogue.  The result
// is at least as close to the float64 reference as the FP32 SIMT kernel.
//
// Persistent kernel, one CTA per SM, up to four 128-thread warpgroups per CTA; every warpgroup owns whole
// 128-position tiles end to end, so no warp idles polling on another role's barrier:
//   prefetch : one elected thread issues cp.async.bulk (TMA 1-D) for the contiguous byte span of input rows
//              the warpgroup's NEXT tile needs (double-buffered, mbarrier expect_tx);
//   build    : thread m turns raw u8 HWC bytes into the fp16 patch row m (dp4a channel sums, magic-number
//              int->fp16) in the canonical K-major no-swizzle core-matrix layout; fence.proxy.async; bar.sync;
//   MMA      : the elected thread issues tcgen05.mma hi (+ lo) into the warpgroup's TMEM accumulator and
//              tcgen05.commit's to an mbarrier; other warpgroups of the SM are building / storing meanwhile;
//   epilogue : tcgen05.ld 16 columns = 4 filters x 4 quadrants -> max4 -> ReLU -> scale -> coalesced stores
//              straight into the filter-major feature row.
// Output rows are written exactly once; input bytes are read exactly once (plus tile-edge halos).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pb {
namespace tcx {

constexpr int RAW_BYTES_MAX = 24 * 1024;   // largest staged span supported (bounds S)
constexpr int A_BYTES = 128 * 32;          // 128 rows x 16 fp16
constexpr int MAX_F = 64;                  // 4*F <= 256 accumulator columns
constexpr int MAX_WG = 4;                  // warpgroups per CTA (TMEM: MAX_WG * columns <= 512)
constexpr int MAX_SLOTS = 6;               // raw staging ring depth per warpgroup (HBM latency >> per-tile work)
constexpr int B_BYTES_MAX = 256 * 32;

struct FastDiv {   // n / d for n < 2^31 by multiply-high: q = umulhi(n, mul) >> shift   (d == 1: mul == 0)
  uint32_t mul, shift, d;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{0u, 0u, d};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                                  // l = ceil(log2 d) >= 1
  f.mul = (uint32_t)((((unsigned long long)1 << (31 + l)) + d - 1) / d);   // ceil(2^(31+l) / d) < 2^32
  f.shift = l - 1;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.mul ? (__umulhi(n, f.mul) >> f.shift) : n;
}

struct ExArgs {
  const uint8_t* images;
  float* out;
  int64_t ld;
  int R;          // total pooled positions (rows) = n * PP
  int num_tiles;  // ceil(R / 128)
  int S, PW, PP;
  FastDiv div_pp, div_pw;
  int F, N;       // N = 4 * round_up(F, 4) accumulator columns
  int has_lo;
  int nwg;        // warpgroups per CTA
  int raw_bytes;  // bytes per raw staging slot (multiple of 128)
  int nslots;     // raw staging slots per warpgroup = prefetch distance + 1 (2..MAX_SLOTS)
  int wg_cols;    // TMEM columns per warpgroup accumulator
  int tmem_cols;  // total allocation (power of two)
  int debug;      // PYCNN_B200_EXTRACT_DEBUG bit0: no global stores, bit1: no patch build, bit2: no MMA/epilogue math
  const uint8_t* bpack;  // [hi | lo], each N x 16 fp16 in canonical core-matrix order
  const float* scale;    // per filter: 2^e / 255
};

struct Span {
  uint32_t b0;     // 16-byte aligned global byte offset of the span
  uint32_t bytes;  // multiple of 16
  uint32_t shift;  // byte offset of pixel g0 inside the span
  uint32_t g0;     // first pixel (global pixel index) of the span
};

__device__ __forceinline__ Span tile_span(const ExArgs& a, int t) {
  const uint32_t r0 = (uint32_t)t * 128u;
  const uint32_t r1 = min((uint32_t)a.R, r0 + 128u) - 1u;
  const uint32_t img0 = fdiv(r0, a.div_pp), py0 = fdiv(r0 - img0 * a.PP, a.div_pw);
  const uint32_t img1 = fdiv(r1, a.div_pp), py1 = fdiv(r1 - img1 * a.PP, a.div_pw);
  const uint32_t SS = (uint32_t)a.S * a.S;
  Span s;
  s.g0 = img0 * SS + 2u * py0 * a.S;
  const uint32_t g1 = img1 * SS + (2u * py1 + 4u) * a.S;
  const uint32_t byte0 = 3u * s.g0;
  s.b0 = byte0 & ~15u;
  s.shift = byte0 - s.b0;
  s.bytes = ((3u * g1 + 15u) & ~15u) - s.b0;
  return s;
}

__device__ __forceinline__ uint32_t pack_sums_to_half2(uint32_t s_lo, uint32_t s_hi) {
  // integers < 1024 -> fp16 exactly: bits 0x6400|x encode 1024+x; subtract 1024 with one HADD2
  const uint32_t biased = __byte_perm(s_lo, s_hi, 0x5410) | 0x64006400u;
  const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&biased), __half2half2(__ushort_as_half(0x6400)));
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(128 * MAX_WG, 1) extract_tc_kernel(const __grid_constant__ ExArgs a) {
  // declared with its alignment so that every pointer derived from it stays in the shared address space
  // (a manual (addr+127)&~127 fix-up degrades all accesses to generic LD/ST)
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: [B hi][B lo][scale][barriers + span info][per warpgroup: raw slot 0, raw slot 1, A tile]
  uint8_t* b_hi = smem;
  uint8_t* b_lo = b_hi + B_BYTES_MAX;
  float* s_scale = reinterpret_cast<float*>(b_lo + B_BYTES_MAX);   // MAX_F floats
  constexpr int BPW = MAX_SLOTS + 1;                                  // barriers per warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scale + MAX_F);      // per wg: raw_full[MAX_SLOTS], mma_done
  uint2* span_info = reinterpret_cast<uint2*>(bars + BPW * MAX_WG);   // per wg: [MAX_SLOTS] {g0, shift}
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(span_info + MAX_SLOTS * MAX_WG);
  uint8_t* wg_base = reinterpret_cast<uint8_t*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wg = warp >> 2, m = threadIdx.x & 127;
  const int NS = a.nslots;
  uint64_t* raw_full = bars + BPW * wg;
  uint64_t* mma_done = raw_full + MAX_SLOTS;
  uint2* my_span = span_info + MAX_SLOTS * wg;
  uint8_t* raw = wg_base + (size_t)wg * (NS * a.raw_bytes + A_BYTES);
  uint8_t* a_tile = raw + NS * a.raw_bytes;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.nwg * BPW; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  {  // filter operand + scales: global -> smem (generic proxy), made visible to the tensor core below
    const int b_bytes = a.N * 32;
    const uint4* src = reinterpret_cast<const uint4*>(a.bpack);
    for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x) {
      reinterpret_cast<uint4*>(b_hi)[i] = src[i];
      reinterpret_cast<uint4*>(b_lo)[i] = src[b_bytes / 16 + i];
    }
    for (int i = threadIdx.x; i < MAX_F; i += blockDim.x) s_scale[i] = (i < a.F) ? a.scale[i] : 0.f;
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  pdl_wait();                 // the images (H2D copy or a previous kernel) are complete from here on
  pdl_launch_dependents();
  const uint32_t tmem_acc = *tmem_slot + (uint32_t)(wg * a.wg_cols);
  const uint32_t lane_taddr = tmem_acc + (static_cast<uint32_t>((warp & 3) * 32) << 16);

  const int tile0 = blockIdx.x * a.nwg + wg;
  const int tile_step = gridDim.x * a.nwg;
  const uint32_t SS = (uint32_t)a.S * a.S, S3 = 3u * a.S;
  const uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, a.N, false, false);
  const uint64_t dbh = tc::make_smem_desc(tc::smem_u32(b_hi), 128, 256, tc::UMMA_SWIZZLE_NONE);
  const uint64_t dbl = tc::make_smem_desc(tc::smem_u32(b_lo), 128, 256, tc::UMMA_SWIZZLE_NONE);
  const uint64_t da = tc::make_smem_desc(tc::smem_u32(a_tile), 128, 256, tc::UMMA_SWIZZLE_NONE);
  uint8_t* const arow = a_tile + (m >> 3) * 256 + (m & 7) * 16;
  const int nchunks = a.N >> 4;
  const int bar_id = 1 + wg;

  // the prefetch thread (warp 1 of the warpgroup) and the MMA thread (warp 0) are different warps on purpose:
  // the extra scalar work is spread instead of making one warp the straggler at every bar.sync
  if (m == 32) {   // fill the ring: the first NS-1 tiles of this warpgroup
    for (int j = 0; j < NS - 1; ++j) {
      const int tj = tile0 + j * tile_step;
      if (tj >= a.num_tiles) break;
      const Span sp = tile_span(a, tj);
      my_span[j] = make_uint2(sp.g0, sp.shift);
      tc::mbar_arrive_expect_tx(&raw_full[j], sp.bytes);
      tc::bulk_load(raw + j * a.raw_bytes, a.images + sp.b0, sp.bytes, &raw_full[j]);
    }
  }
  int it = 0, s = 0;           // s = it % NS
  uint32_t ph = 0;             // (it / NS) & 1
  for (int t = tile0; t < a.num_tiles; t += tile_step, ++it) {
    // ---- prefetch tile it+NS-1 into the slot tile it-1 used (everyone finished reading it before the last bar.sync)
    if (m == 32) {
      const int tn = t + (NS - 1) * tile_step;
      if (tn < a.num_tiles) {
        const int sn = (s == 0) ? NS - 1 : s - 1;
        const Span sp = tile_span(a, tn);
        my_span[sn] = make_uint2(sp.g0, sp.shift);
        tc::mbar_arrive_expect_tx(&raw_full[sn], sp.bytes);
        tc::bulk_load(raw + sn * a.raw_bytes, a.images + sp.b0, sp.bytes, &raw_full[sn]);
      }
    }
    // ---- build: thread m -> row m of the A tile
    const uint32_t r = (uint32_t)t * 128u + m;
    const bool valid = r < (uint32_t)a.R;
    uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t pix = 0, img = 0, p = 0;
    if (valid) {   // position decode overlaps the wait for the raw bytes
      img = fdiv(r, a.div_pp);
      p = r - img * a.PP;
      const uint32_t py = fdiv(p, a.div_pw), px = p - py * a.PW;
      pix = img * SS + 2u * py * a.S + 2u * px;
    }
    tc::mbar_wait_sleepy(&raw_full[s], ph);
    if (valid && !(a.debug & 2)) {
      const uint2 info = my_span[s];
      const uint32_t* words = reinterpret_cast<const uint32_t*>(raw + s * a.raw_bytes);
      uint32_t off = info.y + 3u * (pix - info.x);
#pragma unroll
      for (int u = 0; u < 4; ++u, off += S3) {
        const uint32_t w = off >> 2, sh = (off & 3u) * 8u;
        const uint32_t x0 = words[w], x1 = words[w + 1], x2 = words[w + 2], x3 = words[w + 3];
        const uint32_t W0 = __funnelshift_r(x0, x1, sh), W1 = __funnelshift_r(x1, x2, sh),
                       W2 = __funnelshift_r(x2, x3, sh);
        // bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const uint32_t s0 = __dp4a(W0, 0x00010101u, 0u);
        const uint32_t s1 = __dp4a(W0, 0x01000000u, __dp4a(W1, 0x00000101u, 0u));
        const uint32_t s2 = __dp4a(W1, 0x01010000u, __dp4a(W2, 0x00000001u, 0u));
        const uint32_t s3 = __dp4a(W2, 0x01010100u, 0u);
        h[2 * u] = pack_sums_to_half2(s0, s1);
        h[2 * u + 1] = pack_sums_to_half2(s2, s3);
      }
    }
    *reinterpret_cast<uint4*>(arow) = make_uint4(h[0], h[1], h[2], h[3]);        // k = 0..7  (u = 0,1)
    *reinterpret_cast<uint4*>(arow + 128) = make_uint4(h[4], h[5], h[6], h[7]);  // k = 8..15 (u = 2,3)
    tc::fence_proxy_async_smem();       // A rows visible to the tensor core (async proxy)
    tc::fence_before_thread_sync();     // orders the previous tile's tcgen05.ld before the next MMA
    named_bar_sync(bar_id, 128);
    // ---- MMA: one thread for the warpgroup
    if (m == 0) {
      tc::fence_after_thread_sync();
      tc::mma_f16(tmem_acc, da, dbh, idesc, 0u);
      if (a.has_lo) tc::mma_f16(tmem_acc, da, dbl, idesc, 1u);
      tc::mma_commit(mma_done);
    }
    __syncwarp();
    // ---- epilogue: thread m owns pooled position r; lanes of a warp = consecutive positions
    float* o = a.out + (int64_t)img * a.ld + p;
    tc::mbar_wait_sleepy(mma_done, it & 1);
    tc::fence_after_thread_sync();
    // 16 accumulator columns per tcgen05.ld = 4 filters x 4 pool quadrants.  Per filter: two 3-input max
    // (FMNMX3: quadrant max, then ReLU folded in), one FMUL by the per-filter scale (one LDS.128 per chunk),
    // one store.  All chunks but the last hold four real filters, so only the last one checks f < F.
    const int full_chunks = a.F >> 2;
    const bool st_ok = valid && !(a.debug & 1);
#pragma unroll 1
    for (int c = 0; c < full_chunks; ++c) {
      float v[16];
      tc::tmem_ld16(lane_taddr + c * 16, v);
      const float4 sc = *reinterpret_cast<const float4*>(s_scale + 4 * c);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j, o += a.PP) {
        const float mx = fmaxf(fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), v[4 * j + 2]), fmaxf(v[4 * j + 3], 0.f));
        if (st_ok) *o = mx * scv[j];
      }
    }
    if (full_chunks < nchunks) {   // ragged last chunk (F % 4 filters)
      float v[16];
      tc::tmem_ld16(lane_taddr + full_chunks * 16, v);
      const int rem = a.F & 3;
#pragma unroll
      for (int j = 0; j < 3; ++j, o += a.PP) {
        const float mx = fmaxf(fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), v[4 * j + 2]), fmaxf(v[4 * j + 3], 0.f));
        if (st_ok && j < rem) *o = mx * s_scale[4 * full_chunks + j];
      }
    }
    if (++s == NS) { s = 0; ph ^= 1; }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc(*tmem_slot, (uint32_t)a.tmem_cols);
  }
}

constexpr int SMEM_FIXED = 2 * B_BYTES_MAX + MAX_F * 4 + (MAX_SLOTS + 1) * MAX_WG * 8 + MAX_SLOTS * MAX_WG * 8 + 16 + 256;


}  // namespace tcx

// largest raw span (bytes) a 128-position tile can need for image size S: exact maximum over every
// possible tile start offset inside an image (same arithmetic as tcx::tile_span), cached per S
static inline int64_t tc_extract_max_span(int S) {
  static int cached_S = -1;
  static int64_t cached = 0;
  if (S == cached_S) return cached;
  const int64_t PW = (S - 2) / 2, PP = PW * PW, SS = (int64_t)S * S;
  if (PP <= 0) return 1ll << 30;
  int64_t worst = 0;
  for (int64_t start = 0; start < PP; ++start) {     // r0 mod PP; image index offsets cancel
    const int64_t r0 = start, r1 = start + 127;
    const int64_t img0 = r0 / PP, py0 = (r0 % PP) / PW, img1 = r1 / PP, py1 = (r1 % PP) / PW;
    const int64_t g0 = img0 * SS + 2 * py0 * S, g1 = img1 * SS + (2 * py1 + 4) * S;
    worst = std::max(worst, 3 * (g1 - g0) + 32);
  }
  cached_S = S;
  cached = worst;
  return worst;
}

static inline bool tc_extract_supported(const pycnn_b200_ctx* c, int S) {
  return c->F > 0 && c->F <= tcx::MAX_F && c->d_tc_filters != nullptr && S >= 4 &&
         tc_extract_max_span(S) <= tcx::RAW_BYTES_MAX - 32;
}

// Pack flip(k) into the per-quadrant B operand (hi and lo fp16 parts), canonical K-major no-swizzle
// core-matrix order: element (n, k) at (n/8)*256 + (k/8)*128 + (n%8)*16 + (k%8)*2 bytes; n = f*4 + q.
static inline int tc_extract_set_filters(pycnn_b200_ctx* c) {
  if (c->d_tc_filters) { cudaFree(c->d_tc_filters); c->d_tc_filters = nullptr; }
  if (c->d_tc_scale) { cudaFree(c->d_tc_scale); c->d_tc_scale = nullptr; }
  c->tc_ncols = 0;
  if (c->F > tcx::MAX_F) return 0;  // SIMT kernel handles larger banks
  const int F = c->F, Fp = (int)round_up(F, 4), N = 4 * Fp;
  std::vector<uint16_t> pack((size_t)2 * N * 16, 0);
  std::vector<float> scale(F, 1.0f / 255.0f);
  bool has_lo = false;
  for (int f = 0; f < F; ++f) {
    double mx = 0.0;
    for (int i = 0; i < 9; ++i) mx = std::max(mx, std::fabs(c->h_taps[(size_t)f * 9 + i]));
    int e = 0;
    if (mx > 0.0 && std::isfinite(mx)) e = (int)std::floor(std::log2(mx));
    const double sc = std::ldexp(1.0, -e);  // max|tap| * sc in [1, 2)
    scale[f] = (float)(std::ldexp(1.0, e) / 255.0);
    for (int q = 0; q < 4; ++q) {
      const int qy = q >> 1, qx = q & 1, n = f * 4 + q;
      for (int u = 0; u < 4; ++u)
        for (int v = 0; v < 4; ++v) {
          const int aa = u - qy, bb = v - qx;
          double tap = 0.0;
          if (aa >= 0 && aa <= 2 && bb >= 0 && bb <= 2)
            tap = c->h_taps[(size_t)f * 9 + (2 - aa) * 3 + (2 - bb)] * sc;  // true convolution: flipped
          const __half hi = __float2half_rn((float)tap);
          const double rem = tap - (double)__half2float(hi);
          const __half lo = __float2half_rn((float)rem);
          const int k = u * 4 + v;
          const size_t idx = ((size_t)(n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) / 2;
          pack[idx] = __half_as_ushort(hi);
          pack[(size_t)N * 16 + idx] = __half_as_ushort(lo);
          if (__half_as_ushort(lo) & 0x7fff) has_lo = true;
        }
    }
  }
  PB_CUDA(cudaMalloc(&c->d_tc_filters, pack.size() * 2));
  PB_CUDA(cudaMemcpy(c->d_tc_filters, pack.data(), pack.size() * 2, cudaMemcpyHostToDevice));
  PB_CUDA(cudaMalloc(&c->d_tc_scale, sizeof(float) * F));
  PB_CUDA(cudaMemcpy(c->d_tc_scale, scale.data(), sizeof(float) * F, cudaMemcpyHostToDevice));
  c->tc_ncols = has_lo ? N : -N;  // sign carries has_lo
  return 0;
}

static inline int tc_extract_prepare(pycnn_b200_ctx*) {
  PB_CUDA(cudaFuncSetAttribute(tcx::extract_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  return 0;
}

static inline int tc_extract(pycnn_b200_ctx* c, const uint8_t* d_images, int64_t n, int S, float* out, int64_t ld) {
  const int PW = (S - 2) / 2, PP = PW * PW;
  const int N = std::abs(c->tc_ncols);
  PB_CHECK(N > 0 && PP > 0, "tensor-core extractor not configured");
  int wg_cols = 32;
  while (wg_cols < N) wg_cols <<= 1;                       // accumulator columns per warpgroup (power of two)
  const int nwg = std::min(tcx::MAX_WG, 512 / wg_cols);    // TMEM has 512 columns per SM
  const int tmem_cols = nwg * wg_cols;                     // 128..512, a power of two
  const int raw_bytes = (int)round_up(tc_extract_max_span(S) + 32, 128);
  // ring depth: as deep as shared memory allows (HBM latency is several tiles' worth of work), 2..MAX_SLOTS
  const int64_t budget = 216 * 1024 - tcx::SMEM_FIXED - 128;
  int nslots = (int)((budget / nwg - tcx::A_BYTES) / raw_bytes);
  nslots = std::max(2, std::min(nslots, tcx::MAX_SLOTS));
  const size_t smem_bytes = tcx::SMEM_FIXED + (size_t)nwg * ((size_t)nslots * raw_bytes + tcx::A_BYTES) + 128;
  PB_CHECK(smem_bytes <= 220 * 1024, "tensor-core extractor: staging does not fit in shared memory");
  // keep all device-side index arithmetic in 32 bits: bound rows and bytes per launch
  int64_t per_launch = std::max<int64_t>(1, std::min<int64_t>((1ll << 30) / PP, (1ll << 30) / ((int64_t)3 * S * S)));
  if (per_launch >= 16) per_launch = per_launch / 16 * 16;   // keeps every launch's image base 16-byte aligned
  for (int64_t i0 = 0; i0 < n; i0 += per_launch) {
    const int64_t nb = std::min(per_launch, n - i0);
    tcx::ExArgs a;
    a.images = d_images + i0 * (int64_t)S * S * 3;
    a.out = out + i0 * ld;
    a.ld = ld;
    a.R = (int)(nb * PP);
    a.num_tiles = (int)ceil_div(a.R, 128);
    a.S = S; a.PW = PW; a.PP = PP;
    a.div_pp = tcx::make_fastdiv((uint32_t)PP); a.div_pw = tcx::make_fastdiv((uint32_t)PW);
    a.F = c->F; a.N = N; a.has_lo = c->tc_ncols > 0 ? 1 : 0;
    a.nwg = nwg; a.raw_bytes = raw_bytes; a.nslots = nslots; a.wg_cols = wg_cols; a.tmem_cols = tmem_cols;
    { const char* dbg = getenv("PYCNN_B200_EXTRACT_DEBUG"); a.debug = dbg ? atoi(dbg) : 0; }
    a.bpack = static_cast<const uint8_t*>(c->d_tc_filters);
    a.scale = c->d_tc_scale;
    // the image base must be 16-byte aligned for the bulk copies (cudaMalloc gives 256; chunk offsets keep 16
    // when 3*S*S*i0 is a multiple of 16 -- per_launch is rounded to make it so)
    PB_CHECK((reinterpret_cast<uintptr_t>(a.images) & 15) == 0, "image batch is not 16-byte aligned");
    const int grid = (int)std::min<int64_t>(ceil_div(a.num_tiles, nwg), c->num_sms);
    LaunchScope ls(c, KC_EXTRACT);
    PB_CUDA(launch_pdl(c, tcx::extract_tc_kernel, dim3(grid), dim3(128 * nwg), smem_bytes, c->stream, a));
  }
  return 0;
}

}  // namespace pb
