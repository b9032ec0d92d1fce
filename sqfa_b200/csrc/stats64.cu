// float64 variants of hot path 1: the reference follows the dtype of
// `points` (/root/reference/src/sqfa/statistics.py:28, :32-34), its own remedy for NaN / inf distances
// is "use float64" (_optim.py:28,30,139) and its test-suite runs in float64. Same decomposition as the
// float32 path -- bucketed rows (the int64 label bucketing is shared), per-class sums, Gram of the
// centred rows (upper 64 x 64 tiles), epilogue -- on the FP64 pipe (DFMA): there is no float64 tensor
// path worth its complexity on this part, and no split-precision trick is needed. Every reduction
// runs in a fixed order: results are bit-reproducible.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sqfa_b200.h"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

// sums[c][j] (+)= sum_{i in c} (X[i][j] - shift[c][j]); block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256)
class_sums_f64_kernel(const double* __restrict__ X, int64_t ldx, const int32_t* __restrict__ perm,
                      const int64_t* __restrict__ offsets, const double* __restrict__ shift, int D, double* sums,
                      int accumulate) {
  __shared__ double red[8][32];
  const int c = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  const int64_t r0 = offsets[c], r1 = offsets[c + 1];
  double a = 0.0;
  if (j < D) {
    const double sh = shift != nullptr ? shift[(int64_t)c * D + j] : 0.0;
    for (int64_t r = r0 + ty; r < r1; r += 8) a += X[(int64_t)perm[r] * ldx + j] - sh;
  }
  red[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && j < D) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += red[q][tx];
    double* o = sums + (int64_t)c * D + j;
    *o = accumulate ? *o + s : s;
  }
}

__global__ void class_means_f64_kernel(const double* __restrict__ sums, const int64_t* __restrict__ counts,
                                       const double* __restrict__ shift, int D, int C, double* __restrict__ means) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * D) return;
  const int c = (int)(idx / D);
  means[idx] = sums[idx] / (double)counts[c] + (shift != nullptr ? shift[idx] : 0.0);  // 0 / 0 = NaN when empty
}

// gram[c][m0.., n0..] (+)= sum_{i in c} (x_i - s_c)(x_i - s_c)^T for the 64 x 64 tiles with m0 <= n0.
// 256 threads = 16 x 16, a thread owns a 4 x 4 block; 16 samples per stage in shared memory.
constexpr int G64_TS = 64, G64_KC = 16;

__global__ void __launch_bounds__(256)
gram_f64_kernel(const double* __restrict__ X, int64_t ldx, const int32_t* __restrict__ perm,
                const int64_t* __restrict__ offsets, const double* __restrict__ shift, int D, int TT, double* gram,
                int accumulate) {
  __shared__ double As[G64_KC][G64_TS];
  __shared__ double Bs[G64_KC][G64_TS];
  const int c = blockIdx.y;
  // tile index -> (tm <= tn)
  int t = blockIdx.x, tm = 0;
  while (t >= TT - tm) { t -= TT - tm; ++tm; }
  const int tn = tm + t;
  const int m0 = tm * G64_TS, n0 = tn * G64_TS;
  const bool diag = tm == tn;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lk = tid >> 4, lc = (tid & 15) * 4;  // loader: sample lk of the stage, 4 columns from lc
  const int64_t r0 = offsets[c], r1 = offsets[c + 1];
  double shA[4], shB[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    shA[q] = (shift != nullptr && m0 + lc + q < D) ? shift[(int64_t)c * D + m0 + lc + q] : 0.0;
    shB[q] = (shift != nullptr && n0 + lc + q < D) ? shift[(int64_t)c * D + n0 + lc + q] : 0.0;
  }
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t rb = r0; rb < r1; rb += G64_KC) {
    const int64_t r = rb + lk;
    const bool live = r < r1;
    const double* xr = live ? X + (int64_t)perm[r] * ldx : X;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ca = m0 + lc + q, cb = n0 + lc + q;
      As[lk][lc + q] = (live && ca < D) ? xr[ca] - shA[q] : 0.0;
      if (!diag) Bs[lk][lc + q] = (live && cb < D) ? xr[cb] - shB[q] : 0.0;
    }
    __syncthreads();
    const double (*Bp)[G64_TS] = diag ? As : Bs;
#pragma unroll
    for (int k = 0; k < G64_KC; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[k][ty * 4 + q]; b[q] = Bp[k][tx * 4 + q]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= D) continue;
      double* o = gram + ((int64_t)c * D + row) * D + col;
      *o = accumulate ? *o + acc[i][j] : acc[i][j];
    }
  }
}

// cov[c][i][j] = (G[c][min][max] - n d_i d_j) / (n - ddof), d = means - shift; sm = cov + mu mu^T
__global__ void epilogue_f64_kernel(const double* __restrict__ gram, const double* __restrict__ means,
                                    const double* __restrict__ shift, const int64_t* __restrict__ counts, int D, int C,
                                    int ddof, double* __restrict__ cov, double* __restrict__ sm) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * D * D) return;
  const int c = (int)(idx / ((int64_t)D * D));
  const int i = (int)((idx / D) % D), j = (int)(idx % D);
  const double n = (double)counts[c];
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  double g = gram[((int64_t)c * D + lo) * D + hi];
  const double mi = means[(int64_t)c * D + i], mj = means[(int64_t)c * D + j];
  if (shift != nullptr) {
    const double di = mi - shift[(int64_t)c * D + i], dj = mj - shift[(int64_t)c * D + j];
    g -= n * di * dj;
  }
  const double v = g / (n - (double)ddof);
  cov[idx] = v;
  if (sm != nullptr) sm[idx] = v + mi * mj;
}

// OAS (reference statistics.py:84-93): per class tr(S) and sum S_ij^2 in a fixed order, then
// S <- (1 - rho) S + rho tr(S) / D I,  rho = min(1, ((1 - 2/D) tr(S^2) + tr^2) / ((n + 1 - 2/D)(tr(S^2) - tr^2 / D)))
__global__ void __launch_bounds__(256)
oas_f64_kernel(const double* __restrict__ means, const int64_t* __restrict__ counts, int D, double* __restrict__ cov,
               double* __restrict__ sm) {
  __shared__ double red[2][256];
  __shared__ double s_rho, s_mu;
  const int c = blockIdx.x;
  double* S = cov + (int64_t)c * D * D;
  double tr = 0.0, sq = 0.0;
  for (int64_t e = threadIdx.x; e < (int64_t)D * D; e += 256) {
    const double v = S[e];
    sq += v * v;
    if (e / D == e % D) tr += v;
  }
  red[0][threadIdx.x] = tr;
  red[1][threadIdx.x] = sq;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double t = red[0][0], t2 = red[1][0], n = (double)counts[c], d = (double)D;
    double rho = ((1.0 - 2.0 / d) * t2 + t * t) / ((n + 1.0 - 2.0 / d) * (t2 - t * t / d));
    rho = rho < 1.0 ? rho : 1.0;  // python min(1.0, rho): NaN -> 1.0
    s_rho = rho;
    s_mu = t / d;
  }
  __syncthreads();
  const double rho = s_rho, mu = s_mu;
  for (int64_t e = threadIdx.x; e < (int64_t)D * D; e += 256) {
    const int i = (int)(e / D), j = (int)(e % D);
    const double v = (1.0 - rho) * S[e] + (i == j ? rho * mu : 0.0);
    S[e] = v;
    if (sm != nullptr) sm[(int64_t)c * D * D + e] = v + means[(int64_t)c * D + i] * means[(int64_t)c * D + j];
  }
}

}  // namespace

cudaError_t launch_class_sums_f64(const double* X, int64_t ldx, const int32_t* perm, const int64_t* offsets,
                                  const double* shift, int D, int C, double* sums, int accumulate, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  class_sums_f64_kernel<<<dim3((D + 31) / 32, C), 256, 0, st>>>(X, ldx, perm, offsets, shift, D, sums, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_class_means_f64(const double* sums, const int64_t* counts, const double* shift, int D, int C,
                                   double* means, cudaStream_t st) {
  const int64_t total = (int64_t)C * D;
  if (total <= 0) return cudaSuccess;
  class_means_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(sums, counts, shift, D, C, means);
  return cudaGetLastError();
}

cudaError_t launch_class_gram_f64(const double* X, int64_t ldx, const int32_t* perm, const int64_t* offsets,
                                  const double* shift, int D, int C, double* gram, int accumulate, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const int TT = (D + G64_TS - 1) / G64_TS;
  gram_f64_kernel<<<dim3(TT * (TT + 1) / 2, C), 256, 0, st>>>(X, ldx, perm, offsets, shift, D, TT, gram, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_stats_epilogue_f64(const double* gram, const double* means, const double* shift,
                                      const int64_t* counts, int D, int C, int estimator, int ddof, double* cov,
                                      double* sm, cudaStream_t st) {
  const int64_t total = (int64_t)C * D * D;
  if (total <= 0) return cudaSuccess;
  const bool oas = estimator == SQFA_EST_OAS;
  epilogue_f64_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gram, means, shift, counts, D, C, ddof, cov,
                                                                       oas ? nullptr : sm);
  if (oas) oas_f64_kernel<<<C, 256, 0, st>>>(means, counts, D, cov, sm);
  return cudaGetLastError();
}

}  // namespace sqfa
