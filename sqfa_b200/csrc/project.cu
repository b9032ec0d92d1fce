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

constexpr int PS_WARPS = 4;
constexpr int PS_THREADS = PS_WARPS * 32;
constexpr int PS_COLS = PS_WARPS * 128;  // columns per block: one 128-column strip (float4 per lane) per warp
constexpr int PS_MAXROWS = 256;          // rows of S per block (F^T tile in shared memory)

// partial[split][c][f][j] = sum_{i in rows of split} F[f][i] * S[c][i][j]
// Every warp streams its own 128-column strip down the rows of the split (512 contiguous bytes per
// row and warp, R rows in flight per thread), so there is no cross-warp reduction: a thread's
// accumulators are final for its 4 columns. HBM-bound for k <= 8.
template <int KT>
__global__ void __launch_bounds__(PS_THREADS)
project_stream_kernel(const float* __restrict__ S, const float* __restrict__ F, int C, int D, int k, int rows,
                      int vec16, float* __restrict__ partial) {
  __shared__ __align__(16) float Fs[PS_MAXROWS][KT];  // F^T tile: [row i][filter]
  const int c = blockIdx.z, split = blockIdx.y;
  const int i0 = split * rows;
  const int i1 = min(D, i0 + rows);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < rows * KT; idx += PS_THREADS) {
    const int f = idx / rows, ii = idx % rows;  // consecutive threads read consecutive i: coalesced
    Fs[ii][f] = (i0 + ii < i1 && f < k) ? F[(int64_t)f * D + i0 + ii] : 0.f;
  }
  __syncthreads();
  const int col = blockIdx.x * PS_COLS + warp * 128 + 4 * lane;
  if (col >= D) return;
  // D % 4 == 0 and 16-byte aligned S / partial (checked by the launcher): col + 4 <= D and every
  // row start is 16-byte aligned; otherwise scalar accesses
  const bool vec = vec16 != 0;
  // accumulators packed over filter pairs: acc2[f/2][q] = (filter f, filter f+1) of column col+q,
  // so every update is ONE packed-fp32 FMA (fma.rn.f32x2): FMA issue binds this kernel for k >= 16
  float2 acc2[KT / 2][4];
#pragma unroll
  for (int f = 0; f < KT / 2; ++f)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc2[f][q] = make_float2(0.f, 0.f);
  const float* Sc = S + (int64_t)c * D * D + col;
  constexpr int R = (KT >= 32) ? 4 : 8;  // independent 16-byte loads in flight per thread
  for (int ib = i0; ib < i1; ib += R) {
    float4 v[R];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = ib + u;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < i1) {
        const float* p = Sc + (int64_t)i * D;
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
      const float* fr = Fs[min(ib + u, i1 - 1) - i0];  // rows past the end carry v = 0
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
#pragma unroll
  for (int f = 0; f < KT; ++f) {
    if (f < k) {
      const float2* a = acc2[f / 2];
      const float4 t = (f & 1) ? make_float4(a[0].y, a[1].y, a[2].y, a[3].y)
                               : make_float4(a[0].x, a[1].x, a[2].x, a[3].x);
      float* o = partial + (((int64_t)split * C + c) * k + f) * D + col;
      if (vec) {
        *reinterpret_cast<float4*>(o) = t;
      } else {
        o[0] = t.x;
        if (col + 1 < D) o[1] = t.y;
        if (col + 2 < D) o[2] = t.z;
        if (col + 3 < D) o[3] = t.w;
      }
    }
  }
}

// Second stage of the projection, one block per (class, chunk of 128 columns), one warp per filter f:
//   T[c][f][cols]         = sum over the row splits of the partial products   (fixed order)
//   PsiPart[c][chunk][f][g] = T[c][f][cols] . F[g][cols]      (this chunk's share of Psi = T F^T)
//   MuPart[c][chunk][f]     = F[f][cols] . m[c][cols]         (this chunk's share of mu' = F m)
// The chunk partials are summed per class, in chunk order, by psi_reduce_kernel (stand-alone
// projection) or by class_prepare_kernel (closure): deterministic, and parallel over C * D / 128
// blocks instead of one block per row of T.
constexpr int PF_COLS = 128;

__global__ void __launch_bounds__(1024)
project_finish_kernel(const float* __restrict__ partial, const float* __restrict__ F, const float* __restrict__ M,
                      int C, int D, int k, int nsplit, int vec16, float* __restrict__ T,
                      float* __restrict__ PsiPart, float* __restrict__ MuPart) {
  const int c = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
  const int f = threadIdx.x >> 5, lane = threadIdx.x & 31;  // blockDim.x = 32 k
  const int col = chunk * PF_COLS + 4 * lane;
  const int64_t sstride = (int64_t)C * k * D;
  const int64_t row = ((int64_t)c * k + f) * D;
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  const int nv = col >= D ? 0 : (D - col >= 4 ? 4 : D - col);  // valid columns of this thread
  if (vec16) {
    if (nv == 4) {
      const float* pj = partial + row + col;
      int s = 0;
      for (; s + 4 <= nsplit; s += 4) {  // four independent loads in flight, added in order
        const float4 a0 = *reinterpret_cast<const float4*>(pj + (int64_t)s * sstride);
        const float4 a1 = *reinterpret_cast<const float4*>(pj + (int64_t)(s + 1) * sstride);
        const float4 a2 = *reinterpret_cast<const float4*>(pj + (int64_t)(s + 2) * sstride);
        const float4 a3 = *reinterpret_cast<const float4*>(pj + (int64_t)(s + 3) * sstride);
        t.x += a0.x; t.y += a0.y; t.z += a0.z; t.w += a0.w;
        t.x += a1.x; t.y += a1.y; t.z += a1.z; t.w += a1.w;
        t.x += a2.x; t.y += a2.y; t.z += a2.z; t.w += a2.w;
        t.x += a3.x; t.y += a3.y; t.z += a3.z; t.w += a3.w;
      }
      for (; s < nsplit; ++s) {
        const float4 a0 = *reinterpret_cast<const float4*>(pj + (int64_t)s * sstride);
        t.x += a0.x; t.y += a0.y; t.z += a0.z; t.w += a0.w;
      }
      *reinterpret_cast<float4*>(T + row + col) = t;
    }
  } else {
    float tv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < nv; ++q) {
      for (int s = 0; s < nsplit; ++s) tv[q] += partial[(int64_t)s * sstride + row + col + q];
      T[row + col + q] = tv[q];
    }
    t = make_float4(tv[0], tv[1], tv[2], tv[3]);
  }
  auto load4 = [&](const float* p) {  // 4 columns of a row of F or M, zero beyond D
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec16 && nv == 4) {
      w = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      if (nv > 0) w.x = __ldg(p);
      if (nv > 1) w.y = __ldg(p + 1);
      if (nv > 2) w.z = __ldg(p + 2);
      if (nv > 3) w.w = __ldg(p + 3);
    }
    return w;
  };
  float* pp = PsiPart + (((int64_t)c * nchunk + chunk) * k + f) * k;
  for (int g = 0; g < k; ++g) {
    const float4 w = load4(F + (int64_t)g * D + col);
    float a = (t.x * w.x + t.y * w.y) + (t.z * w.z + t.w * w.w);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) pp[g] = a;
  }
  if (M != nullptr) {
    const float4 w = load4(F + (int64_t)f * D + col);
    const float4 mm = load4(M + (int64_t)c * D + col);
    float a = (w.x * mm.x + w.y * mm.y) + (w.z * mm.z + w.w * mm.w);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) MuPart[((int64_t)c * nchunk + chunk) * k + f] = a;
  }
}

// Psi[c] / Mu[c] = sum over the column chunks, in chunk order (stand-alone projection)
__global__ void psi_reduce_kernel(const float* __restrict__ PsiPart, const float* __restrict__ MuPart, int C, int k,
                                  int nchunk, float* __restrict__ Psi, float* __restrict__ Mu) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int per = k * k + k;
  if (idx >= (int64_t)C * per) return;
  const int c = (int)(idx / per), e = (int)(idx % per);
  float a = 0.f;
  if (e < k * k) {
    for (int ch = 0; ch < nchunk; ++ch) a += PsiPart[((int64_t)c * nchunk + ch) * k * k + e];
    Psi[(int64_t)c * k * k + e] = a;
  } else if (MuPart != nullptr) {
    for (int ch = 0; ch < nchunk; ++ch) a += MuPart[((int64_t)c * nchunk + ch) * k + (e - k * k)];
    Mu[(int64_t)c * k + (e - k * k)] = a;
  }
}

// Sphere constraint (reference constraints.py:37): F[f] = W[f] / |W[f]|, inv_norm[f] = 1 / |W[f]|.
// One block per filter, fixed reduction tree.
__global__ void __launch_bounds__(256)
constraint_fwd_kernel(const float* __restrict__ Wraw, int D, float* __restrict__ F, float* __restrict__ inv_norm) {
  __shared__ float red[256];
  const int f = blockIdx.x;
  const float* w = Wraw + (int64_t)f * D;
  float a = 0.f;
  for (int j = threadIdx.x; j < D; j += 256) a += w[j] * w[j];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float nrm = sqrtf(red[0]);
  for (int j = threadIdx.x; j < D; j += 256) F[(int64_t)f * D + j] = w[j] / nrm;
  if (threadIdx.x == 0) inv_norm[f] = 1.f / nrm;
}

// Last kernel of a closure evaluation, one block per filter f:
//   dF[f] = sum of the class-split partials of the projection adjoint (fixed order),
//   the constraint's adjoint: sphere  grad[f] = (dF[f] - (dF[f] . F[f]) F[f]) / |W[f]|
//                             (autograd through constraints.py:37), none: grad[f] = dF[f];
//   rows f < n_fixed get a zero gradient (FixedFilters detaches them, constraints.py:95-141),
//   out[2] = max |grad| over all filters (atomicMax on the float bits; order-independent).
constexpr int CF_THREADS = 1024;

__device__ __forceinline__ float block_reduce_1024(float v, float* red, bool take_max) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = CF_THREADS / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      red[threadIdx.x] = take_max ? fmaxf(red[threadIdx.x], red[threadIdx.x + o]) : red[threadIdx.x] + red[threadIdx.x + o];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(CF_THREADS)
closure_finish_kernel(const float* __restrict__ partial, int D, int k, int csplit, const float* __restrict__ F,
                      const float* __restrict__ inv_norm, int sphere, int n_fixed, float* __restrict__ grad,
                      float* out, float* out_host, unsigned int* ticket) {
  __shared__ float red[CF_THREADS];
  const int f = blockIdx.x;
  float* g = grad + (int64_t)f * D;
  const int64_t sstride = (int64_t)k * D;
  float dot = 0.f;
  for (int j = threadIdx.x; j < D; j += CF_THREADS) {
    const float* p = partial + (int64_t)f * D + j;
    float a = 0.f;
    int s = 0;
    for (; s + 8 <= csplit; s += 8) {  // eight loads in flight, added in split order
      float u[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) u[q] = p[(int64_t)(s + q) * sstride];
#pragma unroll
      for (int q = 0; q < 8; ++q) a += u[q];
    }
    for (; s < csplit; ++s) a += p[(int64_t)s * sstride];
    g[j] = a;  // re-read below by the same thread
    if (sphere) dot += a * F[(int64_t)f * D + j];
  }
  if (sphere) dot = block_reduce_1024(dot, red, false);
  const float inv = sphere ? inv_norm[f] : 1.f;
  float amax = 0.f;
  for (int j = threadIdx.x; j < D; j += CF_THREADS) {
    float v = g[j];
    if (f < n_fixed) v = 0.f;
    else if (sphere) v = (v - dot * F[(int64_t)f * D + j]) * inv;
    g[j] = v;
    amax = fmaxf(amax, fabsf(v));
  }
  amax = block_reduce_1024(amax, red, true);
  // NaN gradients: fmaxf drops NaN, the loss / non-finite counter carry that information
  if (threadIdx.x == 0 && out != nullptr) {
    atomicMax(reinterpret_cast<unsigned int*>(out + 2), __float_as_uint(amax));
    if (out_host != nullptr) {
      // the last block to arrive publishes {loss, #non-finite, max |grad|} in the caller's mapped host
      // memory: the host then needs an event wait and no copy (`ticket` is zeroed by class_prepare's memset)
      __threadfence();
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        __threadfence();
        volatile float* o = out;
        volatile float* h = out_host;
        h[0] = o[0];
        h[1] = o[1];
        h[2] = o[2];
        __threadfence_system();
      }
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

// Z = X F^T for 16-byte aligned rows: a persistent mini-GEMM. A block walks row tiles of 128 rows;
// per 64-column chunk the X tile (128 x 64, padded rows: conflict-free LDS.128) and the F^T tile
// ([column][filter], so four filters of one column are one LDS.128 broadcast) arrive with cp.async
// while the previous chunk is being multiplied (two buffers). Thread (row pair, g) owns KF = KT/4
// filters of TWO rows (every F value loaded from shared memory feeds two FMAs), accumulated as
// packed pairs: one fma.rn.f32x2 per two filters, row and column.
// HBM-bound for k <= 16 (one read of X), about FMA-bound at k = 32.
constexpr int TT_ROWS = 128, TT_COLS = 64, TT_LDX = TT_COLS + 4, TT_THREADS = 256;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int KF>
__global__ void __launch_bounds__(TT_THREADS)
transform_tile_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ F, int64_t n, int D, int k,
                      float* __restrict__ Z) {
  constexpr int KT = 4 * KF;
  extern __shared__ __align__(16) float tsm[];
  float* Xs = tsm;                              // [2][TT_ROWS][TT_LDX]
  float* Fs = tsm + 2 * TT_ROWS * TT_LDX;       // [2][TT_COLS][KT]
  const int tid = threadIdx.x;
  const int rp = tid & 63, fg = tid >> 6;  // rows rp and rp + 64; a warp: 32 consecutive rows, one filter group
  const int nchunks = (D + TT_COLS - 1) / TT_COLS;
  const int64_t ntiles = (n + TT_ROWS - 1) / TT_ROWS;
  const uint32_t xs_u = (uint32_t)__cvta_generic_to_shared(Xs), fs_u = (uint32_t)__cvta_generic_to_shared(Fs);

  auto load_chunk = [&](int64_t r0, int c0, int buf) {
    // X tile: 128 rows x 16 float4; rows past n and columns past D are zero-filled (src_bytes = 0)
    for (int idx = tid; idx < TT_ROWS * (TT_COLS / 4); idx += TT_THREADS) {
      const int r = idx >> 4, c4 = idx & 15;
      const int64_t gr = r0 + r;
      const int gc = c0 + 4 * c4;
      const bool ok = gr < n && gc < D;  // D % 4 == 0 on this path: a float4 is all inside or all outside
      const float* src = ok ? X + gr * ldx + gc : X;
      cp_async_16(xs_u + (uint32_t)(((buf * TT_ROWS + r) * TT_LDX + 4 * c4) * 4), src, ok ? 16 : 0);
    }
    // F^T tile: Fs[c][f] = F[f][c0 + c]
    for (int idx = tid; idx < TT_COLS * KT; idx += TT_THREADS) {
      const int f = idx / TT_COLS, c = idx % TT_COLS;  // consecutive threads read consecutive columns
      const bool ok = f < k && c0 + c < D;
      const float* src = ok ? F + (int64_t)f * D + c0 + c : F;
      cp_async_4(fs_u + (uint32_t)(((buf * TT_COLS + c) * KT + f) * 4), src, ok ? 4 : 0);
    }
    cp_async_commit();
  };

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * TT_ROWS;
    float2 acc0[KF / 2], acc1[KF / 2];
#pragma unroll
    for (int i = 0; i < KF / 2; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
    load_chunk(r0, 0, 0);
    for (int ch = 0; ch < nchunks; ++ch) {
      const int buf = ch & 1;
      if (ch + 1 < nchunks) {
        load_chunk(r0, (ch + 1) * TT_COLS, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* xr0 = Xs + (buf * TT_ROWS + rp) * TT_LDX;
      const float* xr1 = xr0 + 64 * TT_LDX;
      const float* fb = Fs + buf * TT_COLS * KT + fg * KF;
#pragma unroll 4
      for (int c4 = 0; c4 < TT_COLS / 4; ++c4) {
        const float4 xa = *reinterpret_cast<const float4*>(xr0 + 4 * c4);
        const float4 xb = *reinterpret_cast<const float4*>(xr1 + 4 * c4);
        const float va[4] = {xa.x, xa.y, xa.z, xa.w}, vb[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 a2 = make_float2(va[u], va[u]), b2 = make_float2(vb[u], vb[u]);
          const float* fr = fb + (4 * c4 + u) * KT;
          if constexpr (KF >= 4) {
#pragma unroll
            for (int g = 0; g < KF / 4; ++g) {
              const float4 w = *reinterpret_cast<const float4*>(fr + 4 * g);
              const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
              acc0[2 * g] = __ffma2_rn(a2, w01, acc0[2 * g]);
              acc0[2 * g + 1] = __ffma2_rn(a2, w23, acc0[2 * g + 1]);
              acc1[2 * g] = __ffma2_rn(b2, w01, acc1[2 * g]);
              acc1[2 * g + 1] = __ffma2_rn(b2, w23, acc1[2 * g + 1]);
            }
          } else {
            const float2 w = *reinterpret_cast<const float2*>(fr);
            acc0[0] = __ffma2_rn(a2, w, acc0[0]);
            acc1[0] = __ffma2_rn(b2, w, acc1[0]);
          }
        }
      }
      __syncthreads();  // the buffer is refilled two chunks later
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t gr = r0 + rp + 64 * h;
      if (gr < n) {
#pragma unroll
        for (int i = 0; i < KF / 2; ++i) {
          const int f = fg * KF + 2 * i;
          const float2 a = h ? acc1[i] : acc0[i];
          if (f < k) Z[gr * k + f] = a.x;
          if (f + 1 < k) Z[gr * k + f + 1] = a.y;
        }
      }
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
cudaError_t run_project_stream(const float* S, const float* F, int C, int D, int k, int rows, int nsplit,
                               float* partial, cudaStream_t st) {
  dim3 grid((D + PS_COLS - 1) / PS_COLS, nsplit, C);
  const int vec16 =
      (D % 4 == 0 && ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0) ? 1 : 0;
  project_stream_kernel<KT><<<grid, PS_THREADS, 0, st>>>(S, F, C, D, k, rows, vec16, partial);
  return cudaGetLastError();
}

template <int KT>
cudaError_t run_transform(const float* X, int64_t ldx, const float* F, int64_t n, int D, int k, float* Z,
                          cudaStream_t st) {
  const int smem = KT * 512 * (int)sizeof(float);
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(transform_kernel<KT>, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  const int64_t blocks = (n + 63) / 64;
  for (int f0 = 0; f0 < k; f0 += KT)
    transform_kernel<KT><<<(unsigned)blocks, 256, smem, st>>>(X, ldx, F, n, D, k, f0, Z);
  return cudaGetLastError();
}

}  // namespace

// Resident blocks of project_stream_kernel<KT> per device (occupancy query, cached per device)
template <int KT>
static int project_stream_capacity() {
  static int cap[kMaxDevices] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148 * 6;
  if (cap[dev] == 0) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, project_stream_kernel<KT>, PS_THREADS, 0) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm <= 0 || sms <= 0)
      return 148 * 6;
    cap[dev] = per_sm * sms;
  }
  return cap[dev];
}

static int project_capacity(int k) {
  if (k <= 4) return project_stream_capacity<4>();
  if (k <= 8) return project_stream_capacity<8>();
  if (k <= 16) return project_stream_capacity<16>();
  return project_stream_capacity<32>();
}

// Row splits of the streaming pass. The grid is (column blocks x splits x classes); what costs time is
// a last wave of blocks that fills a fraction of the GPU (c4 with one split: 1000 blocks on 888 resident
// slots = 1.13 waves, 56 % wave efficiency) and, against it, the partial products every split writes and
// the finish kernel reads back (2 nsplit k / D of the bytes of S). Choose the split count that maximises
// wave efficiency / (1 + partial traffic), with at least 64 rows per split and at most PS_MAXROWS.
int project_nsplit(int C, int D, int k) {
  const int64_t colblocks = (D + PS_COLS - 1) / PS_COLS;
  const int64_t per_split = colblocks * (C > 0 ? C : 1);
  const int64_t cap = project_capacity(k);
  const int min_split = (D + PS_MAXROWS - 1) / PS_MAXROWS;
  int max_split = D / 64 > min_split ? D / 64 : min_split;
  if (max_split > 256) max_split = 256;
  int best = min_split;
  double best_score = -1.0;
  for (int n = min_split; n <= max_split; ++n) {
    const int64_t blocks = per_split * n;
    const int64_t waves = (blocks + cap - 1) / cap;
    const double eff = (double)blocks / (double)(waves * cap);
    const double score = eff / (1.0 + 2.0 * n * k / (double)D);
    if (score > best_score * 1.001) { best_score = score; best = n; }  // ties: fewer splits
  }
  return best;
}
int project_rows_per_split(int C, int D, int k) {
  const int n = project_nsplit(C, D, k);
  int rows = (D + n - 1) / n;
  rows = (rows + 7) & ~7;
  if (rows > PS_MAXROWS) rows = PS_MAXROWS;
  return rows;
}
static int project_nsplit_actual(int C, int D, int k) {
  const int rows = project_rows_per_split(C, D, k);
  return (D + rows - 1) / rows;
}

static size_t project_partial_floats(int C, int D, int k) {
  return (size_t)project_nsplit_actual(C, D, k) * C * k * D;
}

int project_nchunk(int D) { return (D + PF_COLS - 1) / PF_COLS; }
static size_t al64(size_t n) { return (n + 63) & ~size_t(63); }
size_t project_psipart_floats(int C, int D, int k) { return al64((size_t)C * project_nchunk(D) * k * k); }
size_t project_mupart_floats(int C, int D, int k) { return al64((size_t)C * project_nchunk(D) * k); }
constexpr int PB_MAX_CSPLIT = 256;
int project_bwd_csplit(int C, int D) {
  // enough (column block, class split) blocks to fill the GPU about 8 times over: every block walks its
  // classes serially (one dependent round of loads per class), so short class lists hide latency
  const int colblocks = (D + 127) / 128;
  int csplit = (8 * 148 + colblocks - 1) / colblocks;
  if (csplit > PB_MAX_CSPLIT) csplit = PB_MAX_CSPLIT;
  if (csplit > C) csplit = C > 0 ? C : 1;
  return csplit < 1 ? 1 : csplit;
}

// workspace (floats): [ row-split partials | PsiPart | MuPart ] for the forward, [ class-split partials ] for
// the adjoint (they alias: the two are never live together)
size_t project_workspace_bytes(int C, int D, int k) {
  const size_t fwd = al64(project_partial_floats(C, D, k)) + project_psipart_floats(C, D, k) +
                     project_mupart_floats(C, D, k);
  const size_t bwd = (size_t)PB_MAX_CSPLIT * k * D;
  return (fwd > bwd ? fwd : bwd) * sizeof(float);
}

// T and the per-chunk partials of Psi / Mu (the closure reduces them inside class_prepare_kernel)
cudaError_t launch_project_partials(const float* S, const float* M, const float* F, int C, int D, int k, float* T,
                                    float* partial, float* PsiPart, float* MuPart, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const int rows = project_rows_per_split(C, D, k);
  int nsplit = project_nsplit_actual(C, D, k);
  cudaError_t e;
  if (project_tc_applicable(S, F, D, k)) {  // k > 8: tensor cores (the SIMT pass is FP32-FMA bound there)
    e = launch_project_tc(S, F, C, D, k, partial, st);
    nsplit = 1;  // the finish kernel finds T itself in the first "row split"
  } else if (k <= 4) e = run_project_stream<4>(S, F, C, D, k, rows, nsplit, partial, st);
  else if (k <= 8) e = run_project_stream<8>(S, F, C, D, k, rows, nsplit, partial, st);
  else if (k <= 16) e = run_project_stream<16>(S, F, C, D, k, rows, nsplit, partial, st);
  else e = run_project_stream<32>(S, F, C, D, k, rows, nsplit, partial, st);
  if (e != cudaSuccess) return e;
  const int vec16 = (D % 4 == 0 && ((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(T) |
                                     reinterpret_cast<uintptr_t>(F) | reinterpret_cast<uintptr_t>(M)) & 15) == 0)
                        ? 1 : 0;
  project_finish_kernel<<<dim3(project_nchunk(D), C), 32 * k, 0, st>>>(partial, F, M, C, D, k, nsplit, vec16, T,
                                                                        PsiPart, MuPart);
  return cudaGetLastError();
}

cudaError_t launch_project_fwd(const float* S, const float* M, const float* F, int C, int D, int k, float* T,
                               float* Psi, float* Mu, float* ws, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  float* PsiPart = ws + al64(project_partial_floats(C, D, k));
  float* MuPart = PsiPart + project_psipart_floats(C, D, k);
  cudaError_t e = launch_project_partials(S, M, F, C, D, k, T, ws, PsiPart, M != nullptr ? MuPart : nullptr, st);
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)C * (k * k + k);
  psi_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(PsiPart, M != nullptr ? MuPart : nullptr, C, k,
                                                                     project_nchunk(D), Psi, Mu);
  return cudaGetLastError();
}

static cudaError_t launch_project_bwd_partials(const float* gPsi, const float* gMu, const float* T, const float* M,
                                               int C, int D, int k, float* ws, int csplit, cudaStream_t st) {
  const int colblocks = (D + 127) / 128;
  const int smem = (k * k + k) * (int)sizeof(float);
  project_bwd_kernel<<<dim3(colblocks, csplit), 128, smem, st>>>(gPsi, gMu, T, M, C, D, k, csplit, ws);
  return cudaGetLastError();
}

cudaError_t launch_project_bwd(const float* gPsi, const float* gMu, const float* T, const float* M, int C, int D,
                               int k, float* dF, float* ws, cudaStream_t st) {
  const int csplit = project_bwd_csplit(C, D);
  cudaError_t e = launch_project_bwd_partials(gPsi, gMu, T, M, C, D, k, ws, csplit, st);
  if (e != cudaSuccess) return e;
  const int64_t total = (int64_t)k * D;
  project_bwd_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(ws, D, k, csplit, dF);
  return cudaGetLastError();
}

// adjoint of the projection followed by the adjoint of the filter constraint (closure tail)
cudaError_t launch_project_bwd_constrained(const float* gPsi, const float* gMu, const float* T, const float* M, int C,
                                           int D, int k, const float* F, const float* inv_norm, int sphere,
                                           int n_fixed, float* grad, float* out, float* out_host,
                                           unsigned int* ticket, float* ws, cudaStream_t st) {
  const int csplit = project_bwd_csplit(C, D);
  cudaError_t e = launch_project_bwd_partials(gPsi, gMu, T, M, C, D, k, ws, csplit, st);
  if (e != cudaSuccess) return e;
  closure_finish_kernel<<<k, CF_THREADS, 0, st>>>(ws, D, k, csplit, F, inv_norm, sphere, n_fixed, grad, out, out_host,
                                                  ticket);
  return cudaGetLastError();
}

cudaError_t launch_constraint_fwd(const float* Wraw, int D, int k, float* F, float* inv_norm, cudaStream_t st) {
  constraint_fwd_kernel<<<k, 256, 0, st>>>(Wraw, D, F, inv_norm);
  return cudaGetLastError();
}

template <int KF>
static cudaError_t run_transform_tile(const float* X, int64_t ldx, const float* F, int64_t n, int D, int k, float* Z,
                                      cudaStream_t st) {
  const int smem = (2 * TT_ROWS * TT_LDX + 2 * TT_COLS * 4 * KF) * (int)sizeof(float);
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(transform_tile_kernel<KF>, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t ntiles = (n + TT_ROWS - 1) / TT_ROWS;
  const int64_t grid = ntiles < 2 * (int64_t)sms ? ntiles : 2 * (int64_t)sms;
  transform_tile_kernel<KF><<<(unsigned)grid, TT_THREADS, smem, st>>>(X, ldx, F, n, D, k, Z);
  return cudaGetLastError();
}

cudaError_t launch_transform(const float* X, int64_t ldx, const float* F, int64_t n, int D, int k, float* Z,
                             cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  if (D % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && k <= 32) {
    if (k <= 8) return run_transform_tile<2>(X, ldx, F, n, D, k, Z, st);
    if (k <= 16) return run_transform_tile<4>(X, ldx, F, n, D, k, Z, st);
    return run_transform_tile<8>(X, ldx, F, n, D, k, Z, st);
  }
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
