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
                       int nsplit, float* __restrict__ partial, float* __restrict__ psi_partial) {
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
  // accumulators packed over filter pairs: acc2[f/2][col] = (acc of filter f, acc of filter f+1)
  // so every update is ONE Blackwell packed-fp32 FMA (fma.rn.f32x2 / FFMA2): this kernel is
  // FMA-issue-bound, not HBM-bound, once k >= 8.
  float2 acc2[KT / 2][4];
#pragma unroll
  for (int f = 0; f < KT / 2; ++f)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc2[f][q] = make_float2(0.f, 0.f);

  const float* Sc = S + (int64_t)c * D * D;
  if (col < D) {
    // R independent 16-byte loads in flight per thread before the FMAs that consume them
    constexpr int R = (KT >= 32) ? 4 : 8;
    for (int ib = i0 + warp * R; ib < i1; ib += PJ_WARPS * R) {
      float4 v[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int i = ib + u;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < i1) {
          const float* p = Sc + (int64_t)i * D + col;
          if (vec) {
            v[u] = __ldg(reinterpret_cast<const float4*>(p));
          } else {
            v[u].x = __ldg(p);
            v[u].y = col + 1 < D ? __ldg(p + 1) : 0.f;
            v[u].z = col + 2 < D ? __ldg(p + 2) : 0.f;
            v[u].w = col + 3 < D ? __ldg(p + 3) : 0.f;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int i = ib + u < i1 ? ib + u : i0;  // rows past the end carry v = 0
        const float* fr = Fs[i - i0];
#pragma unroll
        const float2 vx = make_float2(v[u].x, v[u].x), vy = make_float2(v[u].y, v[u].y);
        const float2 vz = make_float2(v[u].z, v[u].z), vw = make_float2(v[u].w, v[u].w);
#pragma unroll
        for (int f = 0; f < KT; f += 4) {
          const float4 w = *reinterpret_cast<const float4*>(fr + f);
          const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
          acc2[f / 2][0] = __ffma2_rn(w01, vx, acc2[f / 2][0]);
          acc2[f / 2][1] = __ffma2_rn(w01, vy, acc2[f / 2][1]);
          acc2[f / 2][2] = __ffma2_rn(w01, vz, acc2[f / 2][2]);
          acc2[f / 2][3] = __ffma2_rn(w01, vw, acc2[f / 2][3]);
          acc2[f / 2 + 1][0] = __ffma2_rn(w23, vx, acc2[f / 2 + 1][0]);
          acc2[f / 2 + 1][1] = __ffma2_rn(w23, vy, acc2[f / 2 + 1][1]);
          acc2[f / 2 + 1][2] = __ffma2_rn(w23, vz, acc2[f / 2 + 1][2]);
          acc2[f / 2 + 1][3] = __ffma2_rn(w23, vw, acc2[f / 2 + 1][3]);
        }
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
        const float2* a = acc2[f / 2];
        if (f & 1) { t.x += a[0].y; t.y += a[1].y; t.z += a[2].y; t.w += a[3].y; }
        else       { t.x += a[0].x; t.y += a[1].x; t.z += a[2].x; t.w += a[3].x; }
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
  // Psi = T F^T is linear in T: this block adds  sum_{j in its 128 columns} Tpart[f][j] F[g][j].
  // The F^T tile of the main loop is dead now; its memory holds F[:, j0:j0+128].
  float* Fj = &Fs[0][0];  // [KT][128]
  for (int idx = tid; idx < KT * PJ_COLS; idx += PJ_THREADS) {
    const int g = idx / PJ_COLS, jj = idx % PJ_COLS;
    Fj[idx] = (g < k && j0 + jj < D) ? F[(int64_t)g * D + j0 + jj] : 0.f;
  }
  __syncthreads();
  const int nblk = nsplit * gridDim.x;
  const int blk = split * gridDim.x + blockIdx.x;
  float* pp = psi_partial + (int64_t)c * k * k * nblk + blk;  // layout [c][o][blk]
  for (int o = tid; o < k * k; o += PJ_THREADS) {
    const int f = o / k, g = o % k;
    float a = 0.f;
#pragma unroll 8
    for (int jj = 0; jj < PJ_COLS; ++jj) {
      const int j = (jj + lane) & (PJ_COLS - 1);  // per-lane rotation: conflict-free without padding
      a += red[f][j] * Fj[g * PJ_COLS + j];
    }
    pp[(int64_t)o * nblk] = a;
  }
}

// T = sum over row splits of the partial products (elementwise, fully parallel)
__global__ void project_reduce_T_kernel(const float* __restrict__ partial, int64_t total, int nsplit,
                                        float* __restrict__ T) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += partial[(int64_t)s * total + idx];
  T[idx] = a;
}

// Per class: Psi = sum of the per-block partial Psi, mu' = F m (one warp per filter).
__global__ void __launch_bounds__(256)
project_reduce_psi_kernel(const float* __restrict__ psi_partial, const float* __restrict__ F,
                          const float* __restrict__ M, int D, int k, int nblk, float* __restrict__ Psi,
                          float* __restrict__ Mu) {
  const int c = blockIdx.x;
  const float* pp = psi_partial + (int64_t)c * nblk * k * k;  // [o][blk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < k * k; o += 8) {  // one warp per output, lanes stride the blocks
    float a = 0.f;
    for (int b = lane; b < nblk; b += 32) a += pp[(int64_t)o * nblk + b];
    for (int sh = 16; sh > 0; sh >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sh);
    if (lane == 0) Psi[(int64_t)c * k * k + o] = a;
  }
  if (M != nullptr) {
    for (int f = warp; f < k; f += 8) {
      float a = 0.f;
      for (int j = lane; j < D; j += 32) a += F[(int64_t)f * D + j] * M[(int64_t)c * D + j];
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) Mu[(int64_t)c * k + f] = a;
    }
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
                        float* psi_partial, cudaStream_t st) {
  dim3 grid((D + PJ_COLS - 1) / PJ_COLS, nsplit, C);  // k <= KT: a single filter chunk
  project_partial_kernel<KT><<<grid, PJ_THREADS, 0, st>>>(S, F, C, D, k, 0, nsplit, partial, psi_partial);
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

static size_t project_partial_floats(int C, int D, int k) { return (size_t)project_nsplit(D) * C * k * D; }

size_t project_workspace_bytes(int C, int D, int k) {
  const size_t nblk = (size_t)project_nsplit(D) * ((D + PJ_COLS - 1) / PJ_COLS);
  const size_t fwd = (project_partial_floats(C, D, k) + (size_t)C * nblk * k * k) * sizeof(float);
  const size_t bwd = (size_t)64 * k * D * sizeof(float);
  return fwd > bwd ? fwd : bwd;
}

cudaError_t launch_project_fwd(const float* S, const float* M, const float* F, int C, int D, int k, float* T,
                               float* Psi, float* Mu, float* ws, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const int nsplit = project_nsplit(D);
  float* psi_partial = ws + project_partial_floats(C, D, k);
  const int nblk = nsplit * ((D + PJ_COLS - 1) / PJ_COLS);
  cudaError_t e;
  if (k <= 4) e = run_partial<4>(S, F, C, D, k, nsplit, ws, psi_partial, st);
  else if (k <= 8) e = run_partial<8>(S, F, C, D, k, nsplit, ws, psi_partial, st);
  else if (k <= 16) e = run_partial<16>(S, F, C, D, k, nsplit, ws, psi_partial, st);
  else e = run_partial<32>(S, F, C, D, k, nsplit, ws, psi_partial, st);
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)C * k * D;
  project_reduce_T_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, total, nsplit, T);
  project_reduce_psi_kernel<<<C, 256, 0, st>>>(psi_partial, F, M, D, k, nblk, Psi, Mu);
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
