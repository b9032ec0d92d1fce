// K5: per-class factorisation and the pairwise SPD distances with their analytic gradient.
//
// Reference path replaced (/root/reference/src/sqfa/):
//   spd_inv_sqrt            linalg.py:144-162   eigh-based whitening of every class matrix
//   generalized_eigenvalues linalg.py:48-70     C x C conjugations + C*C LAPACK eigvalsh calls
//   affine_invariant(_sq)   distances.py:46-89  sum log^2(lambda), sqrt(. + 1e-6)
//   fisher_rao_lower_bound(_sq) distances.py:177-237  AI^2 of the embedded matrices / 2
//   log_euclidean(_sq), spd_log  distances.py:92-138, linalg.py:165-183
//   closure loss + NaN/inf guard + autograd backward   _optim.py:16-30, 90-96
//
// Formulation. For a pair (i, j) the generalized eigenvalues of (E_i, E_j) are the squared singular
// values of B = L_j^-1 L_i (E = L L^T Cholesky). One warp runs a one-sided (Hestenes) Jacobi on
// the columns of A = B^T held in shared memory, one column per lane (two for 32 < m <= 64),
// round-robin pairing, until all columns are mutually orthogonal: A_f = A V, |a_q|^2 = lambda_q.
// The generalized eigenvectors come for free as Y = L_i^-T A_f (y_q^T E_j y_q = 1), so
//   d(d^2)/dE_i =  sum_q (2 log(lambda_q) / lambda_q) y_q y_q^T
//   d(d^2)/dE_j = -sum_q (2 log(lambda_q))            y_q y_q^T
// -- no 1/(lambda_a - lambda_b) terms (the eigh backward of the reference has them and NaNs on
// repeated eigenvalues). Only the strict lower triangle (i > j) is evaluated: the reference
// computes all C*C pairs and then reads the lower triangle (_optim.py:94).
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sqfa_b200.h"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int PAIR_WARPS = 4;
constexpr float JACOBI_TOL = 1e-6f;  // |a_p . a_q| <= tol |a_p||a_q| counts as orthogonal
constexpr int JACOBI_MAX_SWEEPS = 24;
constexpr float JACOBI_LAST = 3e-4f;  // a sweep of rotations all below this is the last (they leave ~1e-7)
constexpr float DIST_EPS = 1e-6f;    // distances.py:29 EPSILON

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// round-robin (circle method) partner of player p in round r, mp players (mp even)
__device__ __forceinline__ int rr_partner(int p, int r, int mp) {
  const int n1 = mp - 1;
  // players i, j < n1 meet in the round with i + j == 2 r (mod n1); the one left over (2 i == 2 r,
  // i.e. i == r because n1 is odd) meets the fixed player n1
  if (p == n1) return r;
  int q = 2 * r - p;  // in (-n1, 2 n1): one conditional correction replaces the modulo
  if (q < 0) q += n1;
  else if (q >= n1) q -= n1;
  return q == p ? n1 : q;
}

// Cholesky E = L L^T and L^-1 of an m x m SPD matrix, one warp. La / Li: shared, row stride ld.
// Returns false (all lanes) if a pivot is not positive / finite.
__device__ bool warp_cholesky_inverse(const float* __restrict__ E, float* La, float* Li, int m, int ld, int lane) {
  for (int idx = lane; idx < m * m; idx += 32) {
    const int r = idx / m, c = idx % m;
    La[r * ld + c] = E[idx];
    Li[r * ld + c] = 0.f;
  }
  __syncwarp();
  bool ok = true;
  for (int j = 0; j < m; ++j) {
    float d = La[j * ld + j];
    for (int k = 0; k < j; ++k) d -= La[j * ld + k] * La[j * ld + k];
    if (!(d > 0.f) || !isfinite(d)) ok = false;
    const float ljj = sqrtf(d);
    __syncwarp();
    for (int i = j + 1 + lane; i < m; i += 32) {
      float v = La[i * ld + j];
      for (int k = 0; k < j; ++k) v -= La[i * ld + k] * La[j * ld + k];
      La[i * ld + j] = v / ljj;
    }
    if (lane == 0) La[j * ld + j] = ljj;
    __syncwarp();
  }
  // inverse by forward substitution, one column per lane
  for (int c = lane; c < m; c += 32) {
    for (int i = 0; i < m; ++i) {
      float s = (i == c) ? 1.f : 0.f;
      for (int k = 0; k < i; ++k) s -= La[i * ld + k] * Li[k * ld + c];
      Li[i * ld + c] = (i < c) ? 0.f : s / La[i * ld + i];
    }
  }
  __syncwarp();
  return ok;
}

// One-sided Jacobi on the columns of the m x mp matrix in `cur` (row stride ld, column q of slot
// t = lane + 32 t). Double-buffered between cur and nxt; returns the buffer holding the result.
__device__ float* warp_jacobi(float* cur, float* nxt, int m, int mp, int ld, int lane) {
  const int nslot = (mp + 31) >> 5;
  for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; ++sweep) {
    bool rotated = false;
    for (int r = 0; r < mp - 1; ++r) {
      for (int t = 0; t < nslot; ++t) {
        const int p = lane + 32 * t;
        if (p < mp) {
          const int q = rr_partner(p, r, mp);
          float aa = 0.f, bb = 0.f, ab = 0.f;
          for (int s = 0; s < m; ++s) {
            const float x = cur[s * ld + p], y = cur[s * ld + q];
            aa += x * x; bb += y * y; ab += x * y;
          }
          // rotation for the ordered pair (lo, hi): x_lo' = c x_lo - s x_hi, x_hi' = s x_lo + c x_hi
          const bool is_lo = p < q;
          const float alpha = is_lo ? aa : bb, beta = is_lo ? bb : aa;
          float cs = 1.f, sn = 0.f;
          const float ab_sq = ab * ab, scale = alpha * beta;
          if (ab_sq > (JACOBI_TOL * JACOBI_TOL) * scale && alpha > 0.f && beta > 0.f) {
            // same rules as the register kernel: approximate reciprocal / square root for the angle,
            // and a sweep whose rotations were all below JACOBI_LAST is the last one
            const float zeta = (beta - alpha) * rcp_approx(2.f * ab);
            const float tt = copysignf(rcp_approx(fabsf(zeta) + sqrt_approx(fmaf(zeta, zeta, 1.f))), zeta);
            cs = rsqrtf(fmaf(tt, tt, 1.f));
            sn = cs * tt;
            rotated = rotated || ab_sq > (JACOBI_LAST * JACOBI_LAST) * scale;
          }
          const float mine = cs, other = is_lo ? -sn : sn;
          for (int s = 0; s < m; ++s) nxt[s * ld + p] = mine * cur[s * ld + p] + other * cur[s * ld + q];
        }
      }
      __syncwarp();
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (!__any_sync(0xffffffffu, rotated)) break;
  }
  return cur;
}

__device__ __forceinline__ void decode_pair(int64_t p, int& i, int& j) {
  long long ii = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
  while (ii * (ii - 1) / 2 > p) --ii;
  while ((ii + 1) * ii / 2 <= p) ++ii;
  i = (int)ii;
  j = (int)(p - ii * (ii - 1) / 2);
}

__device__ __forceinline__ float finish_distance(float d2, int dist, float* dd_dd2) {
  const int base = dist & 15;
  const float cfac = (base == SQFA_DIST_FISHER_RAO_LB) ? 0.5f : 1.f;
  if (dist & SQFA_DIST_SQUARED) {
    *dd_dd2 = cfac;
    return cfac * d2;
  }
  const float d = sqrtf(cfac * d2 + DIST_EPS);
  *dd_dd2 = cfac / (2.f * d);
  return d;
}

// ------------------------------------------------------------------------------------------------
// per-class factorisation
//   AI / FR: W[c] = [L (m*m) | L^-1 (m*m)]
//   LE     : W[c] = [V (m*m) | lambda (m) | log lambda (m) | logE (m*m)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_WARPS * 32)
class_factor_kernel(const float* __restrict__ E, int C, int m, int dist, float* __restrict__ W,
                    int32_t* __restrict__ flag) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  if (c >= C) return;
  const int mp = (m + 1) & ~1;
  const int ld = (mp > 32 ? 64 : 32) + 1;
  const int per_warp = 4 * m * ld + 2 * m;
  float* La = smem + (size_t)warp * per_warp;
  float* Li = La + m * ld;
  float* bufA = Li + m * ld;
  float* bufB = bufA + m * ld;
  float* lam = bufB + m * ld;
  float* loglam = lam + m;
  const bool ok = warp_cholesky_inverse(E + (int64_t)c * m * m, La, Li, m, ld, lane);
  if (!ok && lane == 0) atomicOr(flag, 1);
  const bool le = (dist & 15) == SQFA_DIST_LOG_EUCLIDEAN;
  if (!le) {
    float* Wc = W + (int64_t)c * 2 * m * m;
    for (int idx = lane; idx < m * m; idx += 32) {
      const int r = idx / m, q = idx % m;
      Wc[idx] = (q <= r) ? La[r * ld + q] : 0.f;
      Wc[m * m + idx] = Li[r * ld + q];
    }
    return;
  }
  // LE: eigendecomposition of E = L L^T through Jacobi on the columns of A0 = L^T
  for (int s = 0; s < m; ++s)
    for (int q = lane; q < (ld - 1); q += 32) bufA[s * ld + q] = (q < m && s <= q) ? La[q * ld + s] : 0.f;
  __syncwarp();
  float* Af = warp_jacobi(bufA, bufB, m, mp, ld, lane);
  float* Vb = (Af == bufA) ? bufB : bufA;
  for (int q = lane; q < m; q += 32) {
    float n2 = 0.f;
    for (int s = 0; s < m; ++s) n2 += Af[s * ld + q] * Af[s * ld + q];
    lam[q] = n2;
    loglam[q] = logf(n2);
    // V[:, q] = L^-T a_q : A_f = L^T V with V orthogonal, so this is already a unit eigenvector
    for (int r = 0; r < m; ++r) {
      float v = 0.f;
      for (int s = r; s < m; ++s) v += Li[s * ld + r] * Af[s * ld + q];
      Vb[r * ld + q] = v;
    }
  }
  __syncwarp();
  float* Wc = W + (int64_t)c * (2 * m * m + 2 * m);
  for (int idx = lane; idx < m * m; idx += 32) {
    const int r = idx / m, s = idx % m;
    Wc[idx] = Vb[r * ld + s];
    float a = 0.f;
    for (int q = 0; q < m; ++q) a += loglam[q] * Vb[r * ld + q] * Vb[s * ld + q];
    Wc[m * m + 2 * m + idx] = a;
  }
  for (int q = lane; q < m; q += 32) {
    Wc[m * m + q] = lam[q];
    Wc[m * m + m + q] = loglam[q];
  }
}

// ------------------------------------------------------------------------------------------------
// pair kernel, affine-invariant family (AI, FR lower bound): one warp per pair
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_ai_kernel(const float* __restrict__ Wa, const float* __restrict__ Wb, int nA, int nB, int m, int dist, int tri,
               int64_t pair_begin, int64_t pair_end, float weight, const float* __restrict__ gD,
               float* __restrict__ dist_out, float* __restrict__ loss, float* gEa, float* gEb,
               float* __restrict__ eig_out) {
  extern __shared__ float smem[];
  __shared__ float s_d[PAIR_WARPS];
  __shared__ float s_bad[PAIR_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p = pair_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const bool active = p < pair_end;
  const int mp = (m + 1) & ~1;
  const int ld = (mp > 32 ? 64 : 32) + 1;
  const int nslot = (mp + 31) >> 5;
  const int per_warp = 2 * m * ld + m * m + 3 * m;
  float* bufA = smem + (size_t)warp * per_warp;
  float* bufB = bufA + m * ld;
  float* Ls = bufB + m * ld;   // m x m, row stride m: L_i, later L_i^-1, later Zj
  float* ci = Ls + m * m;      // per-eigenvalue coefficients
  float* cj = ci + m;
  float* lamv = cj + m;
  float dval = 0.f, bad = 0.f;
  if (active) {
    int i, j;
    if (tri) decode_pair(p, i, j);
    else { i = (int)(p / nB); j = (int)(p % nB); }
    const float* Wi = Wa + (int64_t)i * 2 * m * m;
    const float* Wj = Wb + (int64_t)j * 2 * m * m;
    // stage L_i (row-major) and L_j^-1 transposed: bufB[r][q] = Linv_j[q][r]
    for (int idx = lane; idx < m * m; idx += 32) {
      Ls[idx] = Wi[idx];
      const int q = idx / m, r = idx % m;
      bufB[r * ld + q] = Wj[m * m + idx];
    }
    __syncwarp();
    // A[s][q] = B[q][s] = sum_{r >= s} Linv_j[q][r] L_i[r][s];  columns >= m are zero (dummy players)
    for (int t = 0; t < nslot; ++t) {
      const int q = lane + 32 * t;
      if (q < ld - 1) {
        for (int s = 0; s < m; ++s) {
          float a = 0.f;
          if (q < m)
            for (int r = s; r < m; ++r) a += bufB[r * ld + q] * Ls[r * m + s];
          bufA[s * ld + q] = a;
        }
      }
    }
    __syncwarp();
    float* Af = warp_jacobi(bufA, bufB, m, mp, ld, lane);
    float* Yb = (Af == bufA) ? bufB : bufA;
    // eigenvalues and the distance
    float d2 = 0.f;
    for (int t = 0; t < nslot; ++t) {
      const int q = lane + 32 * t;
      if (q < m) {
        float n2 = 0.f;
        for (int s = 0; s < m; ++s) n2 += Af[s * ld + q] * Af[s * ld + q];
        const float ll = logf(n2);
        d2 += ll * ll;
        ci[q] = 2.f * ll / n2;
        cj[q] = -2.f * ll;
        lamv[q] = n2;
      }
    }
    d2 = warp_sum(d2);
    if (eig_out != nullptr) {  // generalized eigenvalues, descending (linalg.py:69-70)
      __syncwarp();
      float* eo = eig_out + ((int64_t)i * nB + j) * m;
      for (int t = 0; t < nslot; ++t) {
        const int q = lane + 32 * t;
        if (q < m) {
          const float v = lamv[q];
          int rank = 0;
          for (int u = 0; u < m; ++u) rank += (lamv[u] > v || (lamv[u] == v && u < q)) ? 1 : 0;
          eo[rank] = v;
        }
      }
    }
    float dd_dd2;
    dval = finish_distance(d2, dist, &dd_dd2);
    if (!isfinite(dval)) bad = 1.f;
    if (dist_out != nullptr && lane == 0) {
      dist_out[(int64_t)i * nB + j] = dval;
      if (tri) dist_out[(int64_t)j * nB + i] = dval;
    }
    if (gEa != nullptr) {
      float w = weight * dd_dd2;
      if (gD != nullptr) w *= tri ? (gD[(int64_t)i * nB + j] + gD[(int64_t)j * nB + i]) : gD[(int64_t)i * nB + j];
      __syncwarp();
      // Y = L_i^-T A_f : Y[r][q] = sum_{s >= r} Linv_i[s][r] A_f[s][q]
      for (int idx = lane; idx < m * m; idx += 32) Ls[idx] = Wi[m * m + idx];
      __syncwarp();
      for (int t = 0; t < nslot; ++t) {
        const int q = lane + 32 * t;
        if (q < m) {
          for (int r = 0; r < m; ++r) {
            float v = 0.f;
            for (int s = r; s < m; ++s) v += Ls[s * m + r] * Af[s * ld + q];
            Yb[r * ld + q] = v;
          }
        }
      }
      __syncwarp();
      // Zi = Y diag(w ci) -> over A_f's buffer, Zj = Y diag(w cj) -> Ls (row stride m)
      for (int t = 0; t < nslot; ++t) {
        const int q = lane + 32 * t;
        if (q < m) {
          const float a = w * ci[q], b = w * cj[q];
          for (int r = 0; r < m; ++r) {
            const float y = Yb[r * ld + q];
            Af[r * ld + q] = a * y;
            Ls[r * m + q] = b * y;
          }
        }
      }
      __syncwarp();
      // G_i[r][s] = sum_q Zi[r][q] Y[s][q],  G_j[r][s] = sum_q Zj[r][q] Y[s][q];  lane <-> column s
      float* gi = gEa + (int64_t)i * m * m;
      float* gj = gEb + (int64_t)j * m * m;
      for (int t = 0; t < nslot; ++t) {
        const int s = lane + 32 * t;
        if (s < m) {
          for (int r = 0; r < m; ++r) {
            float a = 0.f, b = 0.f;
            for (int q = 0; q < m; ++q) {
              const float y = Yb[s * ld + q];
              a += Af[r * ld + q] * y;
              b += Ls[r * m + q] * y;
            }
            atomicAdd(gi + r * m + s, a);
            atomicAdd(gj + r * m + s, b);
          }
        }
      }
    }
  }
  if (loss != nullptr) {
    if (lane == 0) { s_d[warp] = dval; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_d[w]; b += s_bad[w]; }
      atomicAdd(loss, a);
      if (b != 0.f) atomicAdd(loss + 1, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pair kernel, affine-invariant family, REGISTER-RESIDENT variant for m <= 32 (MP = m rounded up to
// a multiple of 4): lane q keeps column q of A = (L_j^-1 L_i)^T in MP registers; a Jacobi round is
// MP warp shuffles (the partner's column) + 3 MP FMAs; column norms are carried along incrementally
// (alpha' = alpha - t gamma, beta' = beta + t gamma) and recomputed exactly once per sweep.
// Shared memory only stages the triangular factors (broadcast reads) and the final
// sum_q c_q y_q y_q^T. ~3x fewer issue slots per pair than the shared-memory variant above,
// which stays in use for 32 < m <= 64.
// ------------------------------------------------------------------------------------------------
template <int MP, int MJ>  // MP: padded size (multiple of 4, shared-memory strides); MJ: even m, the Jacobi width
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_ai_reg_kernel(const float* __restrict__ Wa, const float* __restrict__ Wb, int nA, int nB, int m, int dist,
                   int tri, int64_t pair_begin, int64_t pair_end, float weight, const float* __restrict__ gD,
                   float* __restrict__ dist_out, float* __restrict__ loss, float* gEa, float* gEb,
                   float* __restrict__ eig_out) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_d[PAIR_WARPS];
  __shared__ float s_bad[PAIR_WARPS];
  constexpr int LDT = 33;
  constexpr int PER_WARP = MP * LDT + 2 * MP * MP + 3;  // sT | sL | sZ (+ pad to keep 16-byte alignment)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t p = pair_begin + (int64_t)blockIdx.x * nwarps + warp;
  float* sT = smem + (size_t)warp * ((PER_WARP + 3) & ~3) + 2 * MP * MP;  // [MP][33]: Linv_j^T, later Y
  float* sL = smem + (size_t)warp * ((PER_WARP + 3) & ~3);               // [MP][MP]: L_i, Linv_i, later Zj
  float* sZ = sL + MP * MP;                                              // [MP][MP]: Zi
  float dval = 0.f, bad = 0.f;
  if (p < pair_end) {
    int i, j;
    if (tri) decode_pair(p, i, j);
    else { i = (int)(p / nB); j = (int)(p % nB); }
    const float* Wi = Wa + (int64_t)i * 2 * m * m;
    const float* Wj = Wb + (int64_t)j * 2 * m * m;
    const int mp = (m + 1) & ~1;
    // ---- stage L_i (row-major, zero padded to MP) and Linv_j transposed
    for (int idx = lane; idx < MP * MP; idx += 32) sL[idx] = 0.f;
    __syncwarp();
    for (int idx = lane; idx < m * m; idx += 32) {
      const int r = idx / m, c = idx % m;
      sL[r * MP + c] = Wi[idx];
      sT[c * LDT + r] = Wj[m * m + idx];  // sT[r'][q] = Linv_j[q][r']
    }
    __syncwarp();
    // ---- column `lane` of A: a[s] = sum_r Linv_j[lane][r] L_i[r][s]   (both factors lower triangular)
    // The column lives in MP / 2 packed register pairs: dot products and rotations are packed-fp32
    // instructions (fma.rn.f32x2 / mul.rn.f32x2), half the FMA issue slots of scalar code.
    float2 a2[MP / 2];
    float y[MP];
#pragma unroll
    for (int s = 0; s < MP / 2; ++s) a2[s] = make_float2(0.f, 0.f);
    if (lane < m) {
#pragma unroll
      for (int r = 0; r < MP; ++r) {
        if (r < m) {
          const float l = sT[r * LDT + lane];
          const float2 l2 = make_float2(l, l);
#pragma unroll
          for (int s4 = 0; s4 < MP / 4; ++s4) {
            const float4 w = *reinterpret_cast<const float4*>(sL + r * MP + 4 * s4);
            a2[2 * s4] = __ffma2_rn(l2, make_float2(w.x, w.y), a2[2 * s4]);
            a2[2 * s4 + 1] = __ffma2_rn(l2, make_float2(w.z, w.w), a2[2 * s4 + 1]);
          }
        }
      }
    }
    // ---- one-sided Jacobi, columns in registers (entries >= MJ are padding zeros and stay zero)
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; ++sweep) {
      float2 n2p = make_float2(0.f, 0.f);
#pragma unroll
      for (int s = 0; s < MJ / 2; ++s) n2p = __ffma2_rn(a2[s], a2[s], n2p);
      float nrm = n2p.x + n2p.y;
      // a sweep whose largest rotation was below JACOBI_LAST leaves off-diagonals of that size
      // squared (quadratic convergence): no verification sweep needed after it
      bool rotated = false;
      for (int r = 0; r < mp - 1; ++r) {
        const int q = lane < mp ? rr_partner(lane, r, mp) : lane;
        const float nq = __shfl_sync(0xffffffffu, nrm, q);
        float2 y2[MJ / 2];
        float2 ab2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int s = 0; s < MJ / 2; ++s) {
          y2[s].x = __shfl_sync(0xffffffffu, a2[s].x, q);
          y2[s].y = __shfl_sync(0xffffffffu, a2[s].y, q);
          ab2 = __ffma2_rn(a2[s], y2[s], ab2);
        }
        const float ab = ab2.x + ab2.y;
        const bool is_lo = lane < q;
        const float alpha = is_lo ? nrm : nq, beta = is_lo ? nq : nrm;
        float cs = 1.f, sn = 0.f;
        const float ab_sq = ab * ab, scale = alpha * beta;
        if (lane < mp && ab_sq > (JACOBI_TOL * JACOBI_TOL) * scale && alpha > 0.f && beta > 0.f) {
          // approximate reciprocal / square root (1 MUFU each): a Jacobi rotation only has to be
          // orthogonal to fp32 precision (cs^2 + sn^2 = 1 from the same rsqrt as before); an angle
          // off by 1e-7 relative leaves an off-diagonal of that size, far below the tolerance
          const float zeta = (beta - alpha) * rcp_approx(2.f * ab);
          const float tt = copysignf(rcp_approx(fabsf(zeta) + sqrt_approx(fmaf(zeta, zeta, 1.f))), zeta);
          cs = rsqrtf(fmaf(tt, tt, 1.f));
          sn = cs * tt;
          nrm = is_lo ? alpha - tt * ab : beta + tt * ab;
          rotated = rotated || ab_sq > (JACOBI_LAST * JACOBI_LAST) * scale;
        }
        const float other = is_lo ? -sn : sn;
        const float2 cs2 = make_float2(cs, cs), ot2 = make_float2(other, other);
#pragma unroll
        for (int s = 0; s < MJ / 2; ++s) a2[s] = __ffma2_rn(cs2, a2[s], __fmul2_rn(ot2, y2[s]));
      }
      if (!__any_sync(0xffffffffu, rotated)) break;
    }
    float a[MP];
#pragma unroll
    for (int s = 0; s < MP / 2; ++s) { a[2 * s] = a2[s].x; a[2 * s + 1] = a2[s].y; }
    // ---- eigenvalues, distance
    float n2 = 0.f;
#pragma unroll
    for (int s = 0; s < MP; ++s) n2 += a[s] * a[s];
    const float ll = lane < m ? logf(n2) : 0.f;
    const float d2 = warp_sum(ll * ll);
    if (eig_out != nullptr) {  // descending order (linalg.py:69-70)
      int rank = 0;
      for (int u = 0; u < m; ++u) {
        const float v = __shfl_sync(0xffffffffu, n2, u);
        rank += (v > n2 || (v == n2 && u < lane)) ? 1 : 0;
      }
      if (lane < m) eig_out[((int64_t)i * nB + j) * m + rank] = n2;
    }
    float dd_dd2;
    dval = finish_distance(d2, dist, &dd_dd2);
    if (!isfinite(dval)) bad = 1.f;
    if (dist_out != nullptr && lane == 0) {
      dist_out[(int64_t)i * nB + j] = dval;
      if (tri) dist_out[(int64_t)j * nB + i] = dval;
    }
    if (gEa != nullptr) {
      float w = weight * dd_dd2;
      if (gD != nullptr) w *= tri ? (gD[(int64_t)i * nB + j] + gD[(int64_t)j * nB + i]) : gD[(int64_t)i * nB + j];
      const float ci = lane < m ? w * 2.f * ll / n2 : 0.f, cj = lane < m ? -w * 2.f * ll : 0.f;
      // ---- Y = L_i^-T A_f : y[r] = sum_{s >= r} Linv_i[s][r] a[s]
      __syncwarp();
      for (int idx = lane; idx < m * m; idx += 32) sL[(idx / m) * MP + idx % m] = Wi[m * m + idx];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < MP; ++r) y[r] = 0.f;
#pragma unroll
      for (int s = 0; s < MP; ++s) {
        if (s < m) {
#pragma unroll
          for (int r4 = 0; r4 < MP / 4; ++r4) {
            const float4 wv = *reinterpret_cast<const float4*>(sL + s * MP + 4 * r4);
            y[4 * r4 + 0] += wv.x * a[s]; y[4 * r4 + 1] += wv.y * a[s]; y[4 * r4 + 2] += wv.z * a[s]; y[4 * r4 + 3] += wv.w * a[s];
          }
        }
      }
      __syncwarp();
      // ---- Y -> sT[r][q], Zi = Y diag(ci) -> sZ[r][q], Zj = Y diag(cj) -> sL[r][q]   (q = lane)
#pragma unroll
      for (int r = 0; r < MP; ++r) {
        const float yr = lane < m ? y[r] : 0.f;
        sT[r * LDT + lane] = yr;
        if (lane < MP) {
          sZ[r * MP + lane] = ci * yr;
          sL[r * MP + lane] = cj * yr;
        }
      }
      __syncwarp();
      // ---- G_i[r][s'] = sum_q Zi[r][q] Y[s'][q], G_j likewise; lane = s', its row of Y in registers
      if (lane < m) {
#pragma unroll
        for (int q = 0; q < MP; ++q) y[q] = sT[lane * LDT + q];
        float* gi = gEa + (int64_t)i * m * m;
        float* gj = gEb + (int64_t)j * m * m;
        for (int r = 0; r < m; ++r) {
          float ga = 0.f, gb = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < MP / 4; ++q4) {
            const float4 zi = *reinterpret_cast<const float4*>(sZ + r * MP + 4 * q4);
            const float4 zj = *reinterpret_cast<const float4*>(sL + r * MP + 4 * q4);
            ga += zi.x * y[4 * q4] + zi.y * y[4 * q4 + 1] + zi.z * y[4 * q4 + 2] + zi.w * y[4 * q4 + 3];
            gb += zj.x * y[4 * q4] + zj.y * y[4 * q4 + 1] + zj.z * y[4 * q4 + 2] + zj.w * y[4 * q4 + 3];
          }
          atomicAdd(gi + r * m + lane, ga);
          atomicAdd(gj + r * m + lane, gb);
        }
      }
    }
  }
  if (loss != nullptr) {
    if (lane == 0) { s_d[warp] = dval; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sa = 0.f, sb = 0.f;
      for (int w = 0; w < nwarps; ++w) { sa += s_d[w]; sb += s_bad[w]; }
      atomicAdd(loss, sa);
      if (sb != 0.f) atomicAdd(loss + 1, sb);
    }
  }
}

template <int MP, int MJ>
static cudaError_t launch_pair_reg(const float* Wa, const float* Wb, int nA, int nB, int m, int dist, int tri,
                                   int64_t pair_begin, int64_t pair_end, float weight, const float* gD,
                                   float* dist_out, float* loss, float* gEa, float* gEb, float* eig_out,
                                   cudaStream_t st) {
  constexpr int per_warp_floats = ((MP * 33 + 2 * MP * MP + 3) + 3) & ~3;
  const int smem = PAIR_WARPS * per_warp_floats * (int)sizeof(float);
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(pair_ai_reg_kernel<MP, MJ>, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  const int64_t npairs = pair_end - pair_begin;
  const unsigned blocks = (unsigned)((npairs + PAIR_WARPS - 1) / PAIR_WARPS);
  pair_ai_reg_kernel<MP, MJ><<<blocks, PAIR_WARPS * 32, smem, st>>>(Wa, Wb, nA, nB, m, dist, tri, pair_begin, pair_end,
                                                                weight, gD, dist_out, loss, gEa, gEb, eig_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// pair kernel, log-Euclidean: d^2 = |logE_i - logE_j|_F^2; gradient w.r.t. the matrix logarithms
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_WARPS * 32)
pair_le_kernel(const float* __restrict__ Wa, const float* __restrict__ Wb, int nA, int nB, int m, int dist, int tri,
               int64_t pair_begin, int64_t pair_end, float weight, const float* __restrict__ gD,
               float* __restrict__ dist_out, float* __restrict__ loss, float* gLa, float* gLb) {
  __shared__ float s_d[PAIR_WARPS];
  __shared__ float s_bad[PAIR_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t p = pair_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int stride = 2 * m * m + 2 * m;
  float dval = 0.f, bad = 0.f;
  if (p < pair_end) {
    int i, j;
    if (tri) decode_pair(p, i, j);
    else { i = (int)(p / nB); j = (int)(p % nB); }
    const float* Li = Wa + (int64_t)i * stride + m * m + 2 * m;
    const float* Lj = Wb + (int64_t)j * stride + m * m + 2 * m;
    float d2 = 0.f;
    for (int idx = lane; idx < m * m; idx += 32) {
      const float t = Li[idx] - Lj[idx];
      d2 += t * t;
    }
    d2 = warp_sum(d2);
    float dd_dd2;
    dval = finish_distance(d2, dist, &dd_dd2);
    if (!isfinite(dval)) bad = 1.f;
    if (dist_out != nullptr && lane == 0) {
      dist_out[(int64_t)i * nB + j] = dval;
      if (tri) dist_out[(int64_t)j * nB + i] = dval;
    }
    if (gLa != nullptr) {
      float w = 2.f * weight * dd_dd2;
      if (gD != nullptr) w *= tri ? (gD[(int64_t)i * nB + j] + gD[(int64_t)j * nB + i]) : gD[(int64_t)i * nB + j];
      for (int idx = lane; idx < m * m; idx += 32) {
        const float t = w * (Li[idx] - Lj[idx]);
        atomicAdd(gLa + (int64_t)i * m * m + idx, t);
        atomicAdd(gLb + (int64_t)j * m * m + idx, -t);
      }
    }
  }
  if (loss != nullptr) {
    if (lane == 0) { s_d[warp] = dval; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_d[w]; b += s_bad[w]; }
      atomicAdd(loss, a);
      if (b != 0.f) atomicAdd(loss + 1, b);
    }
  }
}

// Daleckii-Krein adjoint of logE = V log(Lambda) V^T:  gE += V [ (V^T gLog V) o Gamma ] V^T,
// Gamma_ab = (log l_a - log l_b) / (l_a - l_b), Gamma_aa = 1 / l_a. One warp per class.
__global__ void __launch_bounds__(PAIR_WARPS * 32)
le_factor_bwd_kernel(const float* __restrict__ W, const float* __restrict__ gLog, int C, int m,
                     float* __restrict__ gE) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  if (c >= C) return;
  const int per_warp = 3 * m * m;
  float* V = smem + (size_t)warp * per_warp;
  float* G = V + m * m;
  float* H = G + m * m;
  const float* Wc = W + (int64_t)c * (2 * m * m + 2 * m);
  const float* lam = Wc + m * m;
  const float* ll = Wc + m * m + m;
  for (int idx = lane; idx < m * m; idx += 32) {
    V[idx] = Wc[idx];
    const int r = idx / m, s = idx % m;  // symmetrise the incoming gradient
    G[idx] = 0.5f * (gLog[(int64_t)c * m * m + idx] + gLog[(int64_t)c * m * m + s * m + r]);
  }
  __syncwarp();
  // H = G V
  for (int idx = lane; idx < m * m; idx += 32) {
    const int r = idx / m, b = idx % m;
    float a = 0.f;
    for (int s = 0; s < m; ++s) a += G[r * m + s] * V[s * m + b];
    H[idx] = a;
  }
  __syncwarp();
  // G = (V^T H) o Gamma
  for (int idx = lane; idx < m * m; idx += 32) {
    const int a_ = idx / m, b = idx % m;
    float a = 0.f;
    for (int r = 0; r < m; ++r) a += V[r * m + a_] * H[r * m + b];
    const float la = lam[a_], lb = lam[b];
    const float dl = la - lb;
    float gam;
    if (fabsf(dl) > 1e-4f * fmaxf(la, lb)) gam = (ll[a_] - ll[b]) / dl;
    else gam = 2.f / (la + lb);  // limit of the divided difference of log
    G[idx] = a * gam;
  }
  __syncwarp();
  // H = V G ; gE = H V^T
  for (int idx = lane; idx < m * m; idx += 32) {
    const int r = idx / m, b = idx % m;
    float a = 0.f;
    for (int a_ = 0; a_ < m; ++a_) a += V[r * m + a_] * G[a_ * m + b];
    H[idx] = a;
  }
  __syncwarp();
  for (int idx = lane; idx < m * m; idx += 32) {
    const int r = idx / m, s = idx % m;
    float a = 0.f;
    for (int b = 0; b < m; ++b) a += H[r * m + b] * V[s * m + b];
    gE[(int64_t)c * m * m + idx] += a;
  }
}

__global__ void fill_diagonal_kernel(float* dist_out, int C, float v) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dist_out[(int64_t)c * C + c] = v;
}

}  // namespace

// warps per block so that the per-warp shared-memory scratch fits (large m -> fewer warps)
static int warps_for(int per_warp_bytes) {
  int nw = (200 * 1024) / (per_warp_bytes > 0 ? per_warp_bytes : 1);
  return nw > PAIR_WARPS ? PAIR_WARPS : (nw < 1 ? 1 : nw);
}

size_t class_factor_floats(int m, int dist) {
  return ((dist & 15) == SQFA_DIST_LOG_EUCLIDEAN) ? (size_t)(2 * m * m + 2 * m) : (size_t)(2 * m * m);
}

cudaError_t launch_class_factor(const float* E, int C, int m, int dist, float* W, int32_t* flag, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const int mp = (m + 1) & ~1;
  const int ld = (mp > 32 ? 64 : 32) + 1;
  const int per_warp = (4 * m * ld + 2 * m) * (int)sizeof(float);
  const int nw = warps_for(per_warp);
  const int smem = nw * per_warp;
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(class_factor_kernel, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  class_factor_kernel<<<(C + nw - 1) / nw, nw * 32, smem, st>>>(E, C, m, dist, W, flag);
  return cudaGetLastError();
}

cudaError_t launch_pair_distances(const float* Wa, const float* Wb, int nA, int nB, int m, int dist, int tri,
                                  int64_t pair_begin, int64_t pair_end, float weight, const float* gD,
                                  float* dist_out, float* loss, float* gEa, float* gEb, float* eig_out,
                                  cudaStream_t st) {
  const int64_t npairs = pair_end - pair_begin;
  if (tri && dist_out != nullptr && nA > 0) {
    const float dv = (dist & SQFA_DIST_SQUARED) ? 0.f : sqrtf(DIST_EPS);  // d(i,i): lambda = 1 exactly
    fill_diagonal_kernel<<<(nA + 255) / 256, 256, 0, st>>>(dist_out, nA, dv);
  }
  if (npairs <= 0) return cudaGetLastError();
  if ((dist & 15) == SQFA_DIST_LOG_EUCLIDEAN) {
    const unsigned blocks = (unsigned)((npairs + PAIR_WARPS - 1) / PAIR_WARPS);
    pair_le_kernel<<<blocks, PAIR_WARPS * 32, 0, st>>>(Wa, Wb, nA, nB, m, dist, tri, pair_begin, pair_end, weight, gD,
                                                       dist_out, loss, gEa, gEb);
    return cudaGetLastError();
  }
  if (m <= 32) {  // register-resident Jacobi
#define SQFA_PAIR_REG(MPV)                                                                                   \
  if (((m + 1) & ~1) < MPV)                                                                                    \
    return launch_pair_reg<MPV, MPV - 2>(Wa, Wb, nA, nB, m, dist, tri, pair_begin, pair_end, weight, gD, dist_out, \
                                         loss, gEa, gEb, eig_out, st);                                          \
  return launch_pair_reg<MPV, MPV>(Wa, Wb, nA, nB, m, dist, tri, pair_begin, pair_end, weight, gD, dist_out, loss, \
                                   gEa, gEb, eig_out, st)
    switch ((m + 3) / 4) {
      case 1: SQFA_PAIR_REG(4);
      case 2: SQFA_PAIR_REG(8);
      case 3: SQFA_PAIR_REG(12);
      case 4: SQFA_PAIR_REG(16);
      case 5: SQFA_PAIR_REG(20);
      case 6: SQFA_PAIR_REG(24);
      case 7: SQFA_PAIR_REG(28);
      default: SQFA_PAIR_REG(32);
    }
#undef SQFA_PAIR_REG
  }
  const int mp = (m + 1) & ~1;
  const int ld = (mp > 32 ? 64 : 32) + 1;
  const int per_warp = (2 * m * ld + m * m + 3 * m) * (int)sizeof(float);
  const int nw = warps_for(per_warp);
  const int smem = nw * per_warp;
  const unsigned blocks = (unsigned)((npairs + nw - 1) / nw);
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(pair_ai_kernel, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  pair_ai_kernel<<<blocks, nw * 32, smem, st>>>(Wa, Wb, nA, nB, m, dist, tri, pair_begin, pair_end, weight, gD,
                                                dist_out, loss, gEa, gEb, eig_out);
  return cudaGetLastError();
}

cudaError_t launch_class_factor_bwd(const float* W, const float* gLog, int C, int m, int dist, float* gE,
                                    cudaStream_t st) {
  if ((dist & 15) != SQFA_DIST_LOG_EUCLIDEAN || C <= 0) return cudaSuccess;
  const int per_warp = 3 * m * m * (int)sizeof(float);
  const int nw = warps_for(per_warp);
  const int smem = nw * per_warp;
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(le_factor_bwd_kernel, smem, smem_set);
    if (e != cudaSuccess) return e;
  }
  le_factor_bwd_kernel<<<(C + nw - 1) / nw, nw * 32, smem, st>>>(W, gLog, C, m, gE);
  return cudaGetLastError();
}

}  // namespace sqfa
