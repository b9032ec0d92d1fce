// Pairwise distances between Gaussians that use the MEAN covariance of the pair -- the reference's other
// `distance_fun` plug-ins (/root/reference/src/sqfa/distances.py:240-432):
//   mahalanobis_sq(a, b) = d^T M^-1 d,          d = mu_a - mu_b,  M = (Sigma_a + Sigma_b) / 2      (:283-330)
//   bhattacharyya(a, b)  = maha / 8 + (logdet M - (logdet Sigma_a + logdet Sigma_b) / 2) / 2      (:240-280)
// (mahalanobis, hellinger and fisher_rao_same_cov are scalar maps of these two, applied on the host side).
// The reference materialises (n_a, n_b, k, k) mean covariances and calls torch.linalg.inv / logdet on all
// of them; here one warp owns one pair: Cholesky M = L L^T in shared memory, z = L^-1 d, maha = |z|^2,
// logdet M = 2 sum log L_jj. The backward pass is analytic, with u = M^-1 d:
//   d maha / d mu_a = 2 u,  d maha / d mu_b = -2 u,  d maha / d Sigma_a = d maha / d Sigma_b = -u u^T / 2
//   d logdet M / d Sigma_a = M^-1 / 2,               d logdet Sigma_a / d Sigma_a = Sigma_a^-1
// Per-pair gradient partials are stored and summed per class in a fixed order (no float atomics).
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sqfa_b200.h"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int GP_WARPS = 4;

__device__ __forceinline__ float gp_warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// in-place lower Cholesky of the k x k matrix M (row stride ld), one warp; returns logdet, sets ok
__device__ float gp_cholesky(float* M, int k, int ld, int lane, bool& ok) {
  float logdet = 0.f;
  for (int j = 0; j < k; ++j) {
    float d = M[j * ld + j];
    for (int q = 0; q < j; ++q) d -= M[j * ld + q] * M[j * ld + q];
    if (!(d > 0.f) || !isfinite(d)) ok = false;
    const float ljj = sqrtf(d);
    logdet += 2.f * logf(ljj);
    __syncwarp();
    for (int i = j + 1 + lane; i < k; i += 32) {
      float v = M[i * ld + j];
      for (int q = 0; q < j; ++q) v -= M[i * ld + q] * M[j * ld + q];
      M[i * ld + j] = v / ljj;
    }
    if (lane == 0) M[j * ld + j] = ljj;
    __syncwarp();
  }
  return logdet;
}

// Linv (row stride ld) = L^-1 for lower triangular L, one column per lane
__device__ void gp_tri_inverse(const float* L, float* Li, int k, int ld, int lane) {
  for (int c = lane; c < k; c += 32) {
    for (int i = 0; i < k; ++i) {
      float s = (i == c) ? 1.f : 0.f;
      for (int q = c; q < i; ++q) s -= L[i * ld + q] * Li[q * ld + c];
      Li[i * ld + c] = (i < c) ? 0.f : s / L[i * ld + i];
    }
  }
  __syncwarp();
}

// per class: logdet Sigma_c and (optionally) Sigma_c^-1; one warp per class
__global__ void __launch_bounds__(GP_WARPS * 32)
gauss_class_kernel(const float* __restrict__ Sigma, int C, int k, float* __restrict__ logdet,
                   float* __restrict__ Sinv, int32_t* __restrict__ flag) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * GP_WARPS + warp;
  if (c >= C) return;
  const int ld = k + 1;
  float* L = sm + (size_t)warp * 2 * k * ld;
  float* Li = L + k * ld;
  for (int idx = lane; idx < k * k; idx += 32) L[(idx / k) * ld + idx % k] = Sigma[(int64_t)c * k * k + idx];
  __syncwarp();
  bool ok = true;
  const float ldv = gp_cholesky(L, k, ld, lane, ok);
  if (!ok && lane == 0) atomicOr(flag, 1);
  if (lane == 0) logdet[c] = ldv;
  if (Sinv != nullptr) {
    gp_tri_inverse(L, Li, k, ld, lane);
    for (int idx = lane; idx < k * k; idx += 32) {  // Sigma^-1 = L^-T L^-1
      const int r = idx / k, s = idx % k;
      float a = 0.f;
      for (int q = (r > s ? r : s); q < k; ++q) a += Li[q * ld + r] * Li[q * ld + s];
      Sinv[(int64_t)c * k * k + idx] = a;
    }
  }
}

// One warp per pair (a, b), a in [a0, a1), b in [0, nB). mode 0: mahalanobis_sq, 1: bhattacharyya.
// Forward: dist_out[a][b]. Backward (gD != NULL): per-pair partials, PER = k*k + k floats each,
//   rowpart[(a - a0) * nB + b] = gD[a][b] * (d dist / d Sigma_a | d dist / d mu_a)   (without the Sigma_a^-1 term)
//   colpart[(a - a0) * nB + b] = gD[a][b] * (d dist / d Sigma_b | d dist / d mu_b)
__global__ void __launch_bounds__(GP_WARPS * 32)
gauss_pair_kernel(const float* __restrict__ muA, const float* __restrict__ SigA, const float* __restrict__ muB,
                  const float* __restrict__ SigB, const float* __restrict__ ldA, const float* __restrict__ ldB, int nA,
                  int nB, int k, int mode, int a0, int a1, const float* __restrict__ gD, float* __restrict__ dist_out,
                  float* __restrict__ rowpart, float* __restrict__ colpart, int32_t* __restrict__ flag) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * GP_WARPS + warp;
  if (t >= (int64_t)(a1 - a0) * nB) return;
  const int a = a0 + (int)(t / nB), b = (int)(t % nB);
  const int ld = k + 1;
  float* L = sm + (size_t)warp * (2 * k * ld + 3 * k);
  float* Li = L + k * ld;
  float* dvec = Li + k * ld;
  float* z = dvec + k;
  float* u = z + k;
  for (int idx = lane; idx < k * k; idx += 32)
    L[(idx / k) * ld + idx % k] = 0.5f * (SigA[(int64_t)a * k * k + idx] + SigB[(int64_t)b * k * k + idx]);
  for (int i = lane; i < k; i += 32) dvec[i] = muA[(int64_t)a * k + i] - muB[(int64_t)b * k + i];
  __syncwarp();
  bool ok = true;
  const float logdetM = gp_cholesky(L, k, ld, lane, ok);
  if (!ok && lane == 0) atomicOr(flag, 1);
  // z = L^-1 d (forward substitution; the inner sums are warp reductions)
  for (int j = 0; j < k; ++j) {
    float s = 0.f;
    for (int q = lane; q < j; q += 32) s += L[j * ld + q] * z[q];
    s = gp_warp_sum(s);
    if (lane == 0) z[j] = (dvec[j] - s) / L[j * ld + j];
    __syncwarp();
  }
  float maha = 0.f;
  for (int i = lane; i < k; i += 32) maha += z[i] * z[i];
  maha = gp_warp_sum(maha);
  if (dist_out != nullptr && lane == 0) {
    const float v = mode == 0 ? maha : 0.125f * maha + 0.5f * (logdetM - 0.5f * (ldA[a] + ldB[b]));
    dist_out[(int64_t)a * nB + b] = v;
  }
  if (gD == nullptr) return;
  // u = L^-T z (back substitution)
  for (int j = k - 1; j >= 0; --j) {
    float s = 0.f;
    for (int q = j + 1 + lane; q < k; q += 32) s += L[q * ld + j] * u[q];
    s = gp_warp_sum(s);
    if (lane == 0) u[j] = (z[j] - s) / L[j * ld + j];
    __syncwarp();
  }
  const float g = gD[(int64_t)a * nB + b];
  const int per = k * k + k;
  float* rp = rowpart + t * per;
  float* cp = colpart + t * per;
  // coefficients: dist = cm * maha + cl * logdet M (+ class terms handled per class)
  const float cm = mode == 0 ? 1.f : 0.125f, cl = mode == 0 ? 0.f : 0.5f;
  if (cl != 0.f) gp_tri_inverse(L, Li, k, ld, lane);
  for (int idx = lane; idx < k * k; idx += 32) {
    const int r = idx / k, s = idx % k;
    float v = -0.5f * cm * u[r] * u[s];
    if (cl != 0.f) {  // + cl * M^-1 / 2
      float minv = 0.f;
      for (int q = (r > s ? r : s); q < k; ++q) minv += Li[q * ld + r] * Li[q * ld + s];
      v += 0.5f * cl * minv;
    }
    rp[idx] = g * v;
    cp[idx] = g * v;
  }
  for (int i = lane; i < k; i += 32) {
    const float v = 2.f * cm * u[i] * g;
    rp[k * k + i] = v;
    cp[k * k + i] = -v;
  }
}

// gSigA[a] (=|+=) sum_b rowpart[a][b] (+ class term), gSigB[b] += sum_a colpart[a][b] over the chunk
// [a0, a1); block = class, thread = element of (Sigma | mu). Fixed order.
__global__ void __launch_bounds__(256)
gauss_reduce_kernel(const float* __restrict__ rowpart, const float* __restrict__ colpart, const float* __restrict__ gD,
                    const float* __restrict__ SinvA, const float* __restrict__ SinvB, int nA, int nB, int k, int mode,
                    int a0, int a1, float* gSigA, float* gMuA, float* gSigB, float* gMuB) {
  const int per = k * k + k;
  const int c = blockIdx.x;
  for (int e = threadIdx.x; e < per; e += 256) {
    if (c >= a0 && c < a1) {  // row side: class c of set A, complete for this chunk
      float s = 0.f, gsum = 0.f;
      const float* p = rowpart + (int64_t)(c - a0) * nB * per + e;
      for (int b = 0; b < nB; ++b) {
        s += p[(int64_t)b * per];
        if (mode == 1 && e < k * k) gsum += gD[(int64_t)c * nB + b];
      }
      if (mode == 1 && e < k * k) s -= 0.25f * gsum * SinvA[(int64_t)c * k * k + e];  // -(logdet Sigma_a) / 4
      if (e < k * k) gSigA[(int64_t)c * k * k + e] = s;
      else gMuA[(int64_t)c * k + (e - k * k)] = s;
    }
    if (c < nB) {  // column side: class c of set B accumulates over the chunks
      float s = 0.f, gsum = 0.f;
      for (int a = a0; a < a1; ++a) {
        s += colpart[((int64_t)(a - a0) * nB + c) * per + e];
        if (mode == 1 && e < k * k) gsum += gD[(int64_t)a * nB + c];
      }
      if (mode == 1 && e < k * k) s -= 0.25f * gsum * SinvB[(int64_t)c * k * k + e];
      if (e < k * k) gSigB[(int64_t)c * k * k + e] += s;
      else gMuB[(int64_t)c * k + (e - k * k)] += s;
    }
  }
}

}  // namespace

size_t gauss_workspace_floats(int nA, int nB, int k, int want_grad) {
  // [logdet A | logdet B | Sinv A | Sinv B | flag | row partials | column partials (chunk of rows of A)]
  size_t f = (size_t)nA + nB + 64;
  if (want_grad) {
    f += ((size_t)nA + nB) * k * k;
    const size_t per = (size_t)k * k + k;
    size_t rows = (size_t)(16u << 20) / (per * (nB > 0 ? nB : 1));  // partials of a chunk: <= 2 x 64 MB
    if (rows < 1) rows = 1;
    if (rows > (size_t)nA) rows = nA;
    f += 2 * rows * nB * per;
  }
  return f + 256;
}

cudaError_t launch_gauss_pairs(const float* muA, const float* SigA, const float* muB, const float* SigB, int nA,
                               int nB, int k, int mode, const float* gD, float* dist_out, float* gSigA, float* gMuA,
                               float* gSigB, float* gMuB, float* ws, int32_t* flag_out, cudaStream_t st) {
  if (nA <= 0 || nB <= 0) return cudaSuccess;
  const bool grad = gD != nullptr;
  float* ldA = ws;
  float* ldB = ldA + nA;
  int32_t* flag = reinterpret_cast<int32_t*>(ldB + nB);
  float* SinvA = ldB + nB + 64;
  float* SinvB = SinvA + (grad ? (size_t)nA * k * k : 0);
  float* part = SinvB + (grad ? (size_t)nB * k * k : 0);
  const int ld = k + 1;
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  {
    const int smem = GP_WARPS * 2 * k * ld * (int)sizeof(float);
    static int smem_set[kMaxDevices] = {0};
    if ((e = ensure_dynamic_smem(gauss_class_kernel, smem, smem_set)) != cudaSuccess) return e;
    const bool need_inv = grad && mode == 1;
    gauss_class_kernel<<<(nA + GP_WARPS - 1) / GP_WARPS, GP_WARPS * 32, smem, st>>>(SigA, nA, k, ldA,
                                                                                  need_inv ? SinvA : nullptr, flag);
    gauss_class_kernel<<<(nB + GP_WARPS - 1) / GP_WARPS, GP_WARPS * 32, smem, st>>>(SigB, nB, k, ldB,
                                                                                  need_inv ? SinvB : nullptr, flag);
  }
  const int smem = GP_WARPS * (2 * k * ld + 3 * k) * (int)sizeof(float);
  static int smem_set2[kMaxDevices] = {0};
  if ((e = ensure_dynamic_smem(gauss_pair_kernel, smem, smem_set2)) != cudaSuccess) return e;
  if (!grad) {
    const int64_t npairs = (int64_t)nA * nB;
    gauss_pair_kernel<<<(unsigned)((npairs + GP_WARPS - 1) / GP_WARPS), GP_WARPS * 32, smem, st>>>(
        muA, SigA, muB, SigB, ldA, ldB, nA, nB, k, mode, 0, nA, nullptr, dist_out, nullptr, nullptr, flag);
  } else {
    const size_t per = (size_t)k * k + k;
    size_t rows = (size_t)(16u << 20) / (per * nB);
    if (rows < 1) rows = 1;
    if (rows > (size_t)nA) rows = nA;
    float* rowpart = part;
    float* colpart = part + rows * nB * per;
    if ((e = cudaMemsetAsync(gSigB, 0, (size_t)nB * k * k * sizeof(float), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(gMuB, 0, (size_t)nB * k * sizeof(float), st)) != cudaSuccess) return e;
    const int nC = nA > nB ? nA : nB;
    for (int a0 = 0; a0 < nA; a0 += (int)rows) {
      const int a1 = a0 + (int)rows < nA ? a0 + (int)rows : nA;
      const int64_t npairs = (int64_t)(a1 - a0) * nB;
      gauss_pair_kernel<<<(unsigned)((npairs + GP_WARPS - 1) / GP_WARPS), GP_WARPS * 32, smem, st>>>(
          muA, SigA, muB, SigB, ldA, ldB, nA, nB, k, mode, a0, a1, gD, dist_out, rowpart, colpart, flag);
      gauss_reduce_kernel<<<nC, 256, 0, st>>>(rowpart, colpart, gD, SinvA, SinvB, nA, nB, k, mode, a0, a1, gSigA, gMuA,
                                              gSigB, gMuB);
    }
  }
  if (flag_out != nullptr) {
    if ((e = cudaMemcpyAsync(flag_out, flag, sizeof(int32_t), cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

}  // namespace sqfa
