// K4 / K6 / transform: everything in the closure that touches the D-dimensional data space.
//
//   project_fwd : T[c] = F S[c] (k x D), Psi[c] = T[c] F^T, mu'[c] = F m[c]
//                 = conjugate_matrix(S, F)  /root/reference/src/sqfa/linalg.py:41 (via
//                 model.py:187) and transform(means) model.py:236, in ONE streaming pass over the
//                 C*D*D statistics (HBM-bound: D*D*4 bytes per class, k/2 flop per byte).
//   project_bwd : dF = sum_c (gPsi[c] + gPsi[c]^T) T[c] + gMu[c] m[c]^T  -- the analytic adjoint of
//                 the above; S is symmetric so the saved T replaces a second pass over S
//                 (autograd in the reference re-reads S: _optim.py:95).
//   transform   : Z = X F^T  model.py:236 for user data (HBM-bound on X).
//   embed       : feature noise (model.py:216-217, 537-538) and the Calvo-Oller embedding
//                 (distances.py:162-174) with its adjoint.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sqfa_b200.h"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int PJ_THREADS = 256;
constexpr int PJ_WARPS = PJ_THREADS / 32;
constexpr int PJ_COLS = 128;  // columns per block (one float4 per lane)
constexpr int PJ_ROWS = 256;  // rows of S per block

// partial[split][c][f][j] = sum_{i in rows of split} F[f][i] * S[c][i][j]
template <int KT>
__global__ void __launch_bounds__(PJ_THREADS)
project_partial_kernel(const float* __restrict__ S, const float* __restrict__ F, int C, int D, int k, int f0,
                       int nsplit, float* __restrict__ partial) {
  __shared__ __align__(16) float Fs[PJ_ROWS][KT];      // F^T tile: [row i][filter]
  __shared__ __align__(16) float red[KT][PJ_COLS];     // cross-warp reduction buffer
  const int c = blockIdx.z, split = blockIdx.y;
  const int j0 = blockIdx.x * PJ_COLS;
  const int i0 = split * PJ_ROWS;
  const int i1 = min(D, i0 + PJ_ROWS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int idx = tid; idx < PJ_ROWS * KT; idx += PJ_THREADS) {
    const int ii = idx / KT, f = idx % KT;
    Fs[ii][f] = (i0 + ii < i1 && f0 + f < k) ? F[(int64_t)(f0 + f) * D + i0 + ii] : 0.f;
  }
  for (int idx = tid; idx < KT * PJ_COLS; idx += PJ_THREADS) (&red[0][0])[idx] = 0.f;
  __syncthreads();

  const int col = j0 + 4 * lane;
  const bool vec = (D % 4 == 0) && (col + 4 <= D);
  float acc[KT][4];
#pragma unroll
  for (int f = 0; f < KT; ++f) acc[f][0] = acc[f][1] = acc[f][2] = acc[f][3] = 0.f;

  const float* Sc = S + (int64_t)c * D * D;
  if (col < D) {
#pragma unroll 2
    for (int i = i0 + warp; i < i1; i += PJ_WARPS) {
      float4 v;
      const float* p = Sc + (int64_t)i * D + col;
      if (vec) {
        v = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        v.x = __ldg(p);
        v.y = col + 1 < D ? __ldg(p + 1) : 0.f;
        v.z = col + 2 < D ? __ldg(p + 2) : 0.f;
        v.w = col + 3 < D ? __ldg(p + 3) : 0.f;
      }
      const float* fr = Fs[i - i0];
#pragma unroll
      for (int f = 0; f < KT; f += 4) {
        const float4 w = *reinterpret_cast<const float4*>(fr + f);
        acc[f + 0][0] += w.x * v.x; acc[f + 0][1] += w.x * v.y; acc[f + 0][2] += w.x * v.z; acc[f + 0][3] += w.x * v.w;
        if (KT > 1) { acc[f + 1][0] += w.y * v.x; acc[f + 1][1] += w.y * v.y; acc[f + 1][2] += w.y * v.z; acc[f + 1][3] += w.y * v.w; }
        if (KT > 2) { acc[f + 2][0] += w.z * v.x; acc[f + 2][1] += w.z * v.y; acc[f + 2][2] += w.z * v.z; acc[f + 2][3] += w.z * v.w; }
        if (KT > 3) { acc[f + 3][0] += w.w * v.x; acc[f + 3][1] += w.w * v.y; acc[f + 3][2] += w.w * v.z; acc[f + 3][3] += w.w * v.w; }
      }
    }
  }
  // deterministic cross-warp reduction: warps add in turn
  for (int w = 0; w < PJ_WARPS; ++w) {
    if (warp == w) {
#pragma unroll
      for (int f = 0; f < KT; ++f) {
        float4* r = reinterpret_cast<float4*>(&red[f][4 * lane]);
        float4 t = *r;
        t.x += acc[f][0]; t.y += acc[f][1]; t.z += acc[f][2]; t.w += acc[f][3];
        *r = t;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < KT * PJ_COLS; idx += PJ_THREADS) {
    const int f = idx / PJ_COLS, jj = idx % PJ_COLS;
    if (f0 + f < k && j0 + jj < D)
      partial[(((int64_t)split * C + c) * k + f0 + f) * D + j0 + jj] = red[f][jj];
  }
}

// Per class: T = sum_split partial, Psi = T F^T, mu' = F m.
__global__ void __launch_bounds__(256)
project_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ F, const float* __restrict__ M,
                        int C, int D, int k, int nsplit, float* __restrict__ T, float* __restrict__ Psi,
                        float* __restrict__ Mu) {
  extern __shared__ float sm[];  // Ts[64][k+1], Fs[64][k+1], Ms[64]
  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const int ld = k + 1;
  float* Ts = sm;
  float* Fs = sm + 64 * ld;
  float* Ms = Fs + 64 * ld;
  // each thread owns outputs o = tid, tid+256, ... of the k*k (+k) results
  const int nout = k * k + (M != nullptr ? k : 0);
  float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // k <= 32 -> nout <= 1056 -> <= 5 per thread
  for (int j0 = 0; j0 < D; j0 += 64) {
    __syncthreads();
    for (int idx = tid; idx < 64 * k; idx += 256) {
      const int f = idx / 64, jj = idx % 64;
      const int j = j0 + jj;
      float t = 0.f, fv = 0.f;
      if (j < D) {
        for (int s = 0; s < nsplit; ++s) t += partial[(((int64_t)s * C + c) * k + f) * D + j];
        T[((int64_t)c * k + f) * D + j] = t;
        fv = F[(int64_t)f * D + j];
      }
      Ts[jj * ld + f] = t;
      Fs[jj * ld + f] = fv;
    }
    if (M != nullptr)
      for (int jj = tid; jj < 64; jj += 256) Ms[jj] = (j0 + jj < D) ? M[(int64_t)c * D + j0 + jj] : 0.f;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int o = tid + u * 256;
      if (o < k * k) {
        const int f = o / k, g = o % k;
        float a = acc[u];
        for (int jj = 0; jj < 64; ++jj) a += Ts[jj * ld + f] * Fs[jj * ld + g];
        acc[u] = a;
      } else if (o < nout) {
        const int f = o - k * k;
        float a = acc[u];
        for (int jj = 0; jj < 64; ++jj) a += Fs[jj * ld + f] * Ms[jj];
        acc[u] = a;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 5; ++u) {
    const int o = tid + u * 256;
    if (o < k * k) Psi[(int64_t)c * k * k + o] = acc[u];
    else if (o < nout) Mu[(int64_t)c * k + (o - k * k)] = acc[u];
  }
}

// dF[f][j] = sum_c ( sum_g (gPsi[c][f][g] + gPsi[c][g][f]) T[c][g][j] + gMu[c][f] M[c][j] )
// grid: (ceil(D/128), class splits); partial results reduced by project_bwd_finalize_kernel.
__global__ void __launch_bounds__(128)
project_bwd_kernel(const float* __restrict__ gPsi, const float* __restrict__ gMu, const float* __restrict__ T,
                   const float* __restrict__ M, int C, int D, int k, int csplit, float* __restrict__ partial) {
  extern __shared__ float sm[];  // Gs[k][k] symmetrised, gm[k]
  float* Gs = sm;
  float* gm = sm + k * k;
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int c0 = (int)(((int64_t)C * blockIdx.y) / csplit), c1 = (int)(((int64_t)C * (blockIdx.y + 1)) / csplit);
  float acc[32];
#pragma unroll
  for (int f = 0; f < 32; ++f) acc[f] = 0.f;
  for (int c = c0; c < c1; ++c) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < k * k; idx += 128) {
      const int f = idx / k, g = idx % k;
      Gs[idx] = gPsi[(int64_t)c * k * k + f * k + g] + gPsi[(int64_t)c * k * k + g * k + f];
    }
    if (gMu != nullptr)
      for (int f = threadIdx.x; f < k; f += 128) gm[f] = gMu[(int64_t)c * k + f];
    __syncthreads();
    if (j < D) {
      const float mj = (gMu != nullptr) ? M[(int64_t)c * D + j] : 0.f;
      for (int g = 0; g < k; ++g) {
        const float t = T[((int64_t)c * k + g) * D + j];
#pragma unroll
        for (int f = 0; f < 32; ++f)
          if (f < k) acc[f] += Gs[f * k + g] * t;
      }
      if (gMu != nullptr) {
#pragma unroll
        for (int f = 0; f < 32; ++f)
          if (f < k) acc[f] += gm[f] * mj;
      }
    }
  }
  if (j < D) {
#pragma unroll
    for (int f = 0; f < 32; ++f)
      if (f < k) partial[((int64_t)blockIdx.y * k + f) * D + j] = acc[f];
  }
}

__global__ void project_bwd_finalize_kernel(const float* __restrict__ partial, int D, int k, int csplit,
                                            float* __restrict__ dF) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)k * D) return;
  float a = 0.f;
  for (int s = 0; s < csplit; ++s) a += partial[(int64_t)s * k * D + idx];
  dF[idx] = a;
}

// Z[n][f] = sum_j X[n][j] F[f][j]; one warp per row, F chunk in shared memory.
template <int KT>
__global__ void __launch_bounds__(256)
transform_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ F, int64_t n, int D, int k,
                 int f0, float* __restrict__ Z) {
  extern __shared__ float Fs[];  // [KT][512]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t rows_per_block = 64;
  const int64_t r0 = blockIdx.x * rows_per_block;
  float acc[8][KT];  // 8 rows per warp
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int f = 0; f < KT; ++f) acc[r][f] = 0.f;
  for (int j0 = 0; j0 < D; j0 += 512) {
    __syncthreads();
    for (int idx = tid; idx < KT * 512; idx += 256) {
      const int f = idx / 512, jj = idx % 512;
      Fs[idx] = (f0 + f < k && j0 + jj < D) ? F[(int64_t)(f0 + f) * D + j0 + jj] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int64_t row = r0 + warp * 8 + r;
      if (row >= n) break;
      const float* xr = X + row * ldx + j0;
      for (int jj = lane; jj < 512 && j0 + jj < D; jj += 32) {
        const float x = __ldg(xr + jj);
#pragma unroll
        for (int f = 0; f < KT; ++f) acc[r][f] += x * Fs[f * 512 + jj];
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t row = r0 + warp * 8 + r;
    if (row >= n) break;
#pragma unroll
    for (int f = 0; f < KT; ++f) {
      float v = acc[r][f];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && f0 + f < k) Z[row * k + f0 + f] = v;
    }
  }
}

__global__ void embed_fwd_kernel(const float* __restrict__ Psi, const float* __restrict__ Mu, float noise, int C,
                                 int k, int fr, float* __restrict__ E) {
  const int m = fr ? k + 1 : k;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * m * m) return;
  const int c = (int)(idx / (m * m));
  const int r = (int)((idx / m) % m), s = (int)(idx % m);
  float v;
  if (r < k && s < k) {
    v = Psi[(int64_t)c * k * k + r * k + s] + (r == s ? noise : 0.f);
    if (fr) v += Mu[(int64_t)c * k + r] * Mu[(int64_t)c * k + s];
  } else if (r == k && s == k) {
    v = 1.f;
  } else {
    v = Mu[(int64_t)c * k + (r < k ? r : s)];
  }
  E[idx] = v;
}

// gPsi = gE[:k,:k];  gMu[r] = sum_s (gE[r][s] + gE[s][r]) mu[s] + gE[r][k] + gE[k][r]
__global__ void embed_bwd_kernel(const float* __restrict__ gE, const float* __restrict__ Mu, int C, int k, int fr,
                                 float* __restrict__ gPsi, float* __restrict__ gMu) {
  const int m = fr ? k + 1 : k;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)C * k * (k + 1)) return;
  const int c = (int)(idx / (k * (k + 1)));
  const int r = (int)((idx / (k + 1)) % k), s = (int)(idx % (k + 1));
  const float* g = gE + (int64_t)c * m * m;
  if (s < k) {
    gPsi[(int64_t)c * k * k + r * k + s] = g[r * m + s];
  } else if (fr) {
    float a = g[r * m + k] + g[k * m + r];
    for (int t = 0; t < k; ++t) a += (g[r * m + t] + g[t * m + r]) * Mu[(int64_t)c * k + t];
    gMu[(int64_t)c * k + r] = a;
  }
}

template <int KT>
cudaError_t run_partial(const float* S, const float* F, int C, int D, int k, int nsplit, float* partial,
                        cudaStream_t st) {
  dim3 grid((D + PJ_COLS - 1) / PJ_COLS, nsplit, C);
  for (int f0 = 0; f0 < k; f0 += KT)
    project_partial_kernel<KT><<<grid, PJ_THREADS, 0, st>>>(S, F, C, D, k, f0, nsplit, partial);
  return cudaGetLastError();
}

template <int KT>
cudaError_t run_transform(const float* X, int64_t ldx, const float* F, int64_t n, int D, int k, float* Z,
                          cudaStream_t st) {
  const int smem = KT * 512 * (int)sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(transform_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int64_t blocks = (n + 63) / 64;
  for (int f0 = 0; f0 < k; f0 += KT)
    transform_kernel<KT><<<(unsigned)blocks, 256, smem, st>>>(X, ldx, F, n, D, k, f0, Z);
  return cudaGetLastError();
}

}  // namespace

int project_nsplit(int D) { return (D + PJ_ROWS - 1) / PJ_ROWS; }

size_t project_workspace_bytes(int C, int D, int k) {
  const size_t fwd = (size_t)project_nsplit(D) * C * k * D * sizeof(float);
  const size_t bwd = (size_t)64 * k * D * sizeof(float);
  return fwd > bwd ? fwd : bwd;
}

cudaError_t launch_project_fwd(const float* S, const float* M, const float* F, int C, int D, int k, float* T,
                               float* Psi, float* Mu, float* ws, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const int nsplit = project_nsplit(D);
  cudaError_t e;
  if (k <= 4) e = run_partial<4>(S, F, C, D, k, nsplit, ws, st);
  else if (k <= 8) e = run_partial<8>(S, F, C, D, k, nsplit, ws, st);
  else if (k <= 16) e = run_partial<16>(S, F, C, D, k, nsplit, ws, st);
  else e = run_partial<32>(S, F, C, D, k, nsplit, ws, st);
  if (e != cudaSuccess) return e;
  const int smem = (2 * 64 * (k + 1) + 64) * (int)sizeof(float);
  project_finalize_kernel<<<C, 256, smem, st>>>(ws, F, M, C, D, k, nsplit, T, Psi, Mu);
  return cudaGetLastError();
}

cudaError_t launch_project_bwd(const float* gPsi, const float* gMu, const float* T, const float* M, int C, int D,
                               int k, float* dF, float* ws, cudaStream_t st) {
  int csplit = C < 64 ? (C > 0 ? C : 1) : 64;
  // few column blocks -> more class splits are useful; many -> fewer
  const int colblocks = (D + 127) / 128;
  while (csplit > 1 && colblocks * csplit > 2048) csplit >>= 1;
  const int smem = (k * k + k) * (int)sizeof(float);
  project_bwd_kernel<<<dim3(colblocks, csplit), 128, smem, st>>>(gPsi, gMu, T, M, C, D, k, csplit, ws);
  const int64_t total = (int64_t)k * D;
  project_bwd_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, D, k, csplit, dF);
  return cudaGetLastError();
}

cudaError_t launch_transform(const float* X, int64_t ldx, const float* F, int64_t n, int D, int k, float* Z,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (k <= 4) return run_transform<4>(X, ldx, F, n, D, k, Z, st);
  if (k <= 8) return run_transform<8>(X, ldx, F, n, D, k, Z, st);
  return run_transform<16>(X, ldx, F, n, D, k, Z, st);
}

cudaError_t launch_embed_fwd(const float* Psi, const float* Mu, float noise, int C, int k, int fr, float* E,
                             cudaStream_t st) {
  const int m = fr ? k + 1 : k;
  const int64_t total = (int64_t)C * m * m;
  if (total <= 0) return cudaSuccess;
  embed_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Psi, Mu, noise, C, k, fr, E);
  return cudaGetLastError();
}

cudaError_t launch_embed_bwd(const float* gE, const float* Mu, int C, int k, int fr, float* gPsi, float* gMu,
                             cudaStream_t st) {
  const int64_t total = (int64_t)C * k * (k + 1);
  if (total <= 0) return cudaSuccess;
  embed_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gE, Mu, C, k, fr, gPsi, gMu);
  return cudaGetLastError();
}

}  // namespace sqfa
