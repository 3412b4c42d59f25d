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
// Persistent, warp-specialised (320 threads):
//   warp 0      producer : cp.async.bulk (TMA 1-D) of the contiguous byte span of input rows a tile needs
//   warps 1-4   builders : raw u8 HWC -> channel sums (dp4a) -> fp16 patch rows -> A tile in smem
//                          (canonical K-major no-swizzle core-matrix layout), fence.proxy.async
//   warp 5      MMA      : tcgen05.mma hi (+ lo) into a double-buffered TMEM accumulator, tcgen05.commit
//   warps 6-9   epilogue : tcgen05.ld 16 columns = 4 filters x 4 quadrants -> max4 -> ReLU -> scale ->
//                          coalesced stores straight into the filter-major feature row
// Output rows are written exactly once; input bytes are read exactly once (plus tile-edge halos).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pb {
namespace tcx {

constexpr int RAW_BYTES = 24 * 1024;
constexpr int RAW_STAGES = 3;
constexpr int A_BYTES = 128 * 32;  // 128 rows x 16 fp16
constexpr int A_STAGES = 4;
constexpr int MAX_F = 64;          // 4*F <= 256 accumulator columns per stage
constexpr int THREADS = 320;
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
  int stage_cols; // TMEM columns per accumulator stage (tmem_cols / 2)
  int tmem_cols;
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

__global__ void __launch_bounds__(THREADS, 2) extract_tc_kernel(const __grid_constant__ ExArgs a) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 127) & ~uintptr_t(127));
  uint8_t* raw = smem;                                      // RAW_STAGES x RAW_BYTES
  uint8_t* a_ring = raw + RAW_STAGES * RAW_BYTES;           // A_STAGES x A_BYTES
  uint8_t* b_hi = a_ring + A_STAGES * A_BYTES;              // N x 32 B
  uint8_t* b_lo = b_hi + B_BYTES_MAX;
  float* s_scale = reinterpret_cast<float*>(b_lo + B_BYTES_MAX);  // MAX_F floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scale + MAX_F);
  uint64_t* raw_full = bars;                 // [RAW_STAGES]
  uint64_t* raw_empty = raw_full + RAW_STAGES;
  uint64_t* a_full = raw_empty + RAW_STAGES; // [A_STAGES]
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* tmem_full = a_empty + A_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint2* span_info = reinterpret_cast<uint2*>(tmem_slot + 2);     // [RAW_STAGES] {g0, shift} of the staged span

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&raw_empty[s], 128); }
    for (int s = 0; s < A_STAGES; ++s) { tc::mbar_init(&a_full[s], 128); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 5) tc::tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
  {  // filter operand + scales: global -> smem (generic proxy), made visible to the tensor core below
    const int b_bytes = a.N * 32;
    const uint4* src = reinterpret_cast<const uint4*>(a.bpack);
    for (int i = threadIdx.x; i < b_bytes / 16; i += THREADS) {
      reinterpret_cast<uint4*>(b_hi)[i] = src[i];
      reinterpret_cast<uint4*>(b_lo)[i] = src[b_bytes / 16 + i];
    }
    for (int i = threadIdx.x; i < a.F; i += THREADS) s_scale[i] = a.scale[i];
  }
  tc::fence_proxy_async_smem();
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int tile_step = gridDim.x;

  if (warp == 0) {
    // ===== producer: one bulk copy per tile; publishes {g0, shift} of the span next to the data =====
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;   // parity to wait for on raw_empty: first pass falls through
      for (int t = blockIdx.x; t < a.num_tiles; t += tile_step) {
        tc::mbar_wait(&raw_empty[s], ph);
        const Span sp = tile_span(a, t);
        span_info[s] = make_uint2(sp.g0, sp.shift);
        tc::mbar_arrive_expect_tx(&raw_full[s], sp.bytes);     // release: span_info is visible to waiters
        tc::bulk_load(raw + s * RAW_BYTES, a.images + sp.b0, sp.bytes, &raw_full[s]);
        if (++s == RAW_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp <= 4) {
    // ===== builders: thread m builds row m of the A tile =====
    const int m = threadIdx.x - 32;
    int sr = 0, sa = 0;
    uint32_t ph_r = 0, ph_a = 1;
    const uint32_t SS = (uint32_t)a.S * a.S, S3 = 3u * a.S;
    uint8_t* const arow0 = a_ring + (m >> 3) * 256 + (m & 7) * 16;
    for (int t = blockIdx.x; t < a.num_tiles; t += tile_step) {
      const uint32_t r = (uint32_t)t * 128u + m;
      uint32_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      uint32_t pix = 0;
      const bool valid = r < (uint32_t)a.R;
      if (valid) {   // position decode overlaps the wait for the raw bytes
        const uint32_t img = fdiv(r, a.div_pp), p = r - img * a.PP, py = fdiv(p, a.div_pw), px = p - py * a.PW;
        pix = img * SS + 2u * py * a.S + 2u * px;
      }
      tc::mbar_wait(&raw_full[sr], ph_r);
      if (valid) {
        const uint2 info = span_info[sr];
        const uint32_t* words = reinterpret_cast<const uint32_t*>(raw + sr * RAW_BYTES);
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
      tc::mbar_wait(&a_empty[sa], ph_a);
      uint8_t* arow = arow0 + sa * A_BYTES;
      *reinterpret_cast<uint4*>(arow) = make_uint4(h[0], h[1], h[2], h[3]);        // k = 0..7  (u = 0,1)
      *reinterpret_cast<uint4*>(arow + 128) = make_uint4(h[4], h[5], h[6], h[7]);  // k = 8..15 (u = 2,3)
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&a_full[sa]);
      tc::mbar_arrive(&raw_empty[sr]);
      if (++sr == RAW_STAGES) { sr = 0; ph_r ^= 1; }
      if (++sa == A_STAGES) { sa = 0; ph_a ^= 1; }
    }
  } else if (warp == 5) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tc::make_idesc(tc::FMT_F16, 128, a.N, false, false);
      const uint64_t dbh = tc::make_smem_desc(tc::smem_u32(b_hi), 128, 256, tc::UMMA_SWIZZLE_NONE);
      const uint64_t dbl = tc::make_smem_desc(tc::smem_u32(b_lo), 128, 256, tc::UMMA_SWIZZLE_NONE);
      const uint64_t da0 = tc::make_smem_desc(tc::smem_u32(a_ring), 128, 256, tc::UMMA_SWIZZLE_NONE);
      int sa = 0, acc = 0;
      uint32_t ph_a = 0, ph_t = 1;
      for (int t = blockIdx.x; t < a.num_tiles; t += tile_step) {
        tc::mbar_wait(&tmem_empty[acc], ph_t);
        tc::mbar_wait(&a_full[sa], ph_a);
        tc::fence_after_thread_sync();
        const uint64_t da = da0 + (uint64_t)((sa * A_BYTES) >> 4);   // start-address field is in 16-byte units
        const uint32_t d = tmem_base + (uint32_t)(acc * a.stage_cols);
        tc::mma_f16(d, da, dbh, idesc, 0u);
        if (a.has_lo) tc::mma_f16(d, da, dbl, idesc, 1u);
        tc::mma_commit(&a_empty[sa]);
        tc::mma_commit(&tmem_full[acc]);
        if (++sa == A_STAGES) { sa = 0; ph_a ^= 1; }
        if (++acc == 2) { acc = 0; ph_t ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 6..9): TMEM lane quarter = warp % 4 =====
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int acc = 0;
    uint32_t ph_t = 0;
    const int nchunks = a.N >> 4;
    for (int t = blockIdx.x; t < a.num_tiles; t += tile_step) {
      const uint32_t r = (uint32_t)t * 128u + m;
      const bool valid = r < (uint32_t)a.R;
      const uint32_t img = valid ? fdiv(r, a.div_pp) : 0u, p = valid ? r - img * a.PP : 0u;
      float* o = a.out + (int64_t)img * a.ld + p;
      tc::mbar_wait(&tmem_full[acc], ph_t);
      tc::fence_after_thread_sync();
      const uint32_t taddr = lane_base + (uint32_t)(acc * a.stage_cols);
      int f = 0;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        float v[16];
        __syncwarp();
        tc::tmem_ld16(taddr + c * 16, v);
#pragma unroll
        for (int j = 0; j < 4; ++j, ++f, o += a.PP) {
          float mx = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
          mx = fmaxf(mx, 0.f) * s_scale[f];      // s_scale is padded to MAX_F: no bound check on the load
          if (valid && f < a.F) *o = mx;
        }
      }
      __syncwarp();
      tc::fence_before_thread_sync();
      tc::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; ph_t ^= 1; }
    }
  }
  tc::fence_before_thread_sync();
  __syncthreads();
  if (warp == 5) {
    tc::fence_after_thread_sync();
    tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
  }
}

constexpr int SMEM_BYTES = RAW_STAGES * RAW_BYTES + A_STAGES * A_BYTES + 2 * B_BYTES_MAX + MAX_F * 4 + 320 + 128;

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
         tc_extract_max_span(S) <= tcx::RAW_BYTES - 32;
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
  PB_CUDA(cudaFuncSetAttribute(tcx::extract_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tcx::SMEM_BYTES));
  return 0;
}

static inline int tc_extract(pycnn_b200_ctx* c, const uint8_t* d_images, int64_t n, int S, float* out, int64_t ld) {
  const int PW = (S - 2) / 2, PP = PW * PW;
  const int N = std::abs(c->tc_ncols);
  PB_CHECK(N > 0 && PP > 0, "tensor-core extractor not configured");
  int tmem_cols = 32;
  while (tmem_cols < 2 * N) tmem_cols <<= 1;
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
    a.tmem_cols = tmem_cols; a.stage_cols = tmem_cols / 2;
    a.bpack = static_cast<const uint8_t*>(c->d_tc_filters);
    a.scale = c->d_tc_scale;
    // the image base must be 16-byte aligned for the bulk copies (cudaMalloc gives 256; chunk offsets keep 16
    // when 3*S*S*i0 is a multiple of 16 -- per_launch is rounded to make it so)
    PB_CHECK((reinterpret_cast<uintptr_t>(a.images) & 15) == 0, "image batch is not 16-byte aligned");
    const int ctas_per_sm = tmem_cols <= 256 ? 2 : 1;
    const int grid = std::min(a.num_tiles, ctas_per_sm * c->num_sms);
    LaunchScope ls(c, KC_EXTRACT);
    tcx::extract_tc_kernel<<<grid, tcx::THREADS, tcx::SMEM_BYTES, c->stream>>>(a);
    PB_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace pb
