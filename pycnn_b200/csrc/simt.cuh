// simt.cuh -- FP32 CUDA-core implementations: the verification mode the north star asks for
// (1e-5 parity against the float64 reference) and the fallback for shapes/inputs the tensor-core
// kernels do not take.  These are real kernels on the product path when math mode / extract mode
// is *_SIMT; they are never a CPU fallback.
#pragma once
#include "common.cuh"

namespace pb {

// ---------------------------------------------------------------------------------------------
// K1 (SIMT form): fused  sum_c conv3x3_valid(flip(k), ch_c) -> ReLU -> 2x2 max-pool -> flatten.
// pycnn/pycnn.py:273-287 + max_pooling.pyx:21-28.
//  * channels are summed first (linearity) on exact integer values, the /255 of
//    pycnn.py:273 is applied once at the end (ReLU and max commute with a positive scale);
//  * taps arrive already flipped (true convolution): taps[f][a*3+b] multiplies x[i+a][j+b];
//  * one thread = one pooled output for ALL filters: its 4x4 window lives in registers;
//  * output row layout is filter-major: out[n*ld + f*PH*PW + py*PW + px].
// grid (tiles_x*tiles_y, n); block 256 = 16x16 pooled positions; smem: summed tile + taps.
// ---------------------------------------------------------------------------------------------
constexpr int EX_T = 16;               // pooled tile edge
constexpr int EX_IN = 2 * EX_T + 2;    // input tile edge (34)

__global__ void __launch_bounds__(256) extract_simt_kernel(const uint8_t* __restrict__ images, int S, int PH,
                                                           int PW, int F, const float* __restrict__ taps,
                                                           float* __restrict__ out, int64_t ld, int tiles_x) {
  extern __shared__ float smem[];
  pdl_wait();
  float* s_sum = smem;                       // [EX_IN][EX_IN+1]
  float* s_taps = smem + EX_IN * (EX_IN + 1);  // [F][9]
  const int n = blockIdx.y;
  const int ty0 = (blockIdx.x / tiles_x) * EX_T, tx0 = (blockIdx.x % tiles_x) * EX_T;
  const uint8_t* img = images + (int64_t)n * S * S * 3;
  for (int i = threadIdx.x; i < F * 9; i += blockDim.x) s_taps[i] = taps[i];
  // channel-summed input tile (zero outside the image; such pixels only feed cropped outputs)
  for (int i = threadIdx.x; i < EX_IN * EX_IN; i += blockDim.x) {
    const int r = i / EX_IN, c = i - r * EX_IN;
    const int gy = 2 * ty0 + r, gx = 2 * tx0 + c;
    float v = 0.f;
    if (gy < S && gx < S) {
      const uint8_t* p = img + ((int64_t)gy * S + gx) * 3;
      v = (float)((int)p[0] + (int)p[1] + (int)p[2]);
    }
    s_sum[r * (EX_IN + 1) + c] = v;
  }
  __syncthreads();
  const int ly = threadIdx.x / EX_T, lx = threadIdx.x % EX_T;
  const int py = ty0 + ly, px = tx0 + lx;
  if (py >= PH || px >= PW) return;
  float w[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) w[a][b] = s_sum[(2 * ly + a) * (EX_IN + 1) + 2 * lx + b];
  float* o = out + (int64_t)n * ld + (int64_t)py * PW + px;
  const int64_t plane = (int64_t)PH * PW;
  const float inv255 = 1.0f / 255.0f;
  for (int f = 0; f < F; ++f) {
    const float* t = s_taps + f * 9;
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const float k = t[a * 3 + b];
        c00 = fmaf(w[a][b], k, c00);
        c01 = fmaf(w[a][b + 1], k, c01);
        c10 = fmaf(w[a + 1][b], k, c10);
        c11 = fmaf(w[a + 1][b + 1], k, c11);
      }
    const float mx = fmaxf(fmaxf(fmaxf(c00, c01), fmaxf(c10, c11)), 0.f);
    o[f * plane] = mx * inv255;
  }
}

// ---------------------------------------------------------------------------------------------
// K2 (SIMT form): C[M,N] = A . B^T-ish with generic strides, FP32 FMA, fused epilogues.
//   A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
//   epi 0: C = acc          (dW, backward_pass.pyx:71,92)
//   epi 1: C = acc + bias   (output logits, forward_pass.pyx:36-37)
//   epi 2: C = relu(acc+b)  (hidden layers, forward_pass.pyx:29-32)
//   epi 3: C = acc where mask>0 else 0   (dA then ReLU mask, backward_pass.pyx:85-86,97)
// 64x64 tile, BK 16, 256 threads, 4x4 register micro-tile.  Operands are FP32, products and sums are
// FP64 (DFMA): as the *verification* mode its job is to remove accumulation-order noise so that what is
// left against the float64 reference is only the FP32 storage of weights/activations.
// ---------------------------------------------------------------------------------------------
struct SimtGemmArgs {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbn, sbk;
  float* C; int64_t ldc;
  int M, N, K, epi;
  const float* bias;
  const float* mask; int64_t ldmask;
};

__global__ void __launch_bounds__(256) sgemm_simt_kernel(SimtGemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  pdl_wait();
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  double acc[4][4] = {};
  const bool a_kfast = (g.sak == 1), b_kfast = (g.sbk == 1);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int id = tid + i * 256;
      int m, k;
      if (a_kfast) { m = id >> 4; k = id & 15; } else { m = id & 63; k = id >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < g.K) ? g.A[(int64_t)gm * g.sam + (int64_t)gk * g.sak] : 0.f;
      int n;
      if (b_kfast) { n = id >> 4; k = id & 15; } else { n = id & 63; k = id >> 6; }
      const int gn = n0 + n;
      const int gk2 = k0 + k;
      Bs[k][n] = (gn < g.N && gk2 < g.K) ? g.B[(int64_t)gn * g.sbn + (int64_t)gk2 * g.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; b[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma((double)a[i], (double)b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = (g.epi == 1 || g.epi == 2) ? (float)(acc[i][j] + (double)g.bias[n]) : (float)acc[i][j];
      if (g.epi == 2) v = (v < 0.f) ? 0.f : v;   // relu_inplace: NaN stays NaN (forward_pass.pyx:12)
      if (g.epi == 3) v = (g.mask[(int64_t)m * g.ldmask + n] <= 0.f) ? 0.f : v;   // dz[z<=0]=0 (backward_pass.pyx:86)
      g.C[(int64_t)m * g.ldc + n] = v;
    }
  }
}

}  // namespace pb
