// K2: segmented per-class Gram accumulation  G_c = sum_{i in class c} (x_i - s_c)(x_i - s_c)^T
// on the sm_100a tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM) as 3xTF32
// split precision:  x = hi + lo,  G ~= hi^T hi + hi^T lo + lo^T hi   (fp32-level accuracy).
//
// Replaces the per-class loop of the reference (gather `points[indices]`, centre, einsum
// "ij,ik->jk"):  /root/reference/src/sqfa/statistics.py:36-47 and :113-122.
//
// Data flow per CTA (persistent, one CTA per SM, dynamic job counter):
//   8 producer warps : gather class rows through the bucket permutation with coalesced LDG.128
//                      straight into registers (no raw staging in smem -- shared-memory bandwidth
//                      is the binding resource for 3xTF32), subtract the class
//                      shift, split hi/lo, store both into the MN-major 128B-swizzled UMMA
//                      operand layout, fence.proxy.async, arrive on full[stage].
//   1 MMA warp       : one elected lane issues 3 tcgen05.mma per K=8 step (lo*hi, hi*lo, hi*hi),
//                      tcgen05.commit -> empty[stage]; after the last K block commit -> tmem_full.
//   epilogue         : the producer warps read the 128 x N accumulator with tcgen05.ld and store
//                      (or red.add when the class is split along K / accumulating) to gram.
//
// Only tiles that intersect the upper triangle are computed; K3 mirrors them.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int BM = 128;  // output rows per tile  (A operand width, UMMA M)
constexpr int BN = 256;  // output cols per tile  (B operand width, UMMA N max)
constexpr int BK = 16;   // samples per pipeline stage (2 UMMA K-steps of 8)
constexpr int STAGES = 4;
constexpr int PROD_WARPS = 8;
constexpr int GRAM_THREADS = (PROD_WARPS + 1) * 32;

constexpr int CHUNK_BYTES = BK * 128;              // one 32-column chunk: BK rows of 128 B
constexpr int A_BYTES = (BM / 32) * CHUNK_BYTES;   //  8 KB
constexpr int B_BYTES = (BN / 32) * CHUNK_BYTES;   // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // hi+lo for A and B = 48 KB
constexpr int GRAM_SMEM = STAGES * STAGE_BYTES + 1024;  // + alignment slack
constexpr uint32_t TMEM_COLS = 256;

constexpr uint32_t LAYOUT_SW128 = 2;

struct GramParams {
  const float* X;
  int64_t ldx;
  const int32_t* perm;       // rows sorted by class (stable)
  const int64_t* offsets;    // C+1 class offsets into perm
  const float* shift;        // C x D (or nullptr -> 0)
  float* gram;               // C x D x D
  int* job_counter;
  int D;
  int C;
  int TM, TN, T;             // tile grid and tiles per class
  int KS;                    // K splits per tile
  int atomic_out;            // 1 -> red.add into gram, 0 -> plain store
  int vec_ok;                // 16-byte aligned rows -> LDG.128 / STG.128
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Load one float4 column group of one sample row (zero outside [0,D) / invalid row).
__device__ __forceinline__ float4 load_group(const float* row, int col, int D, bool row_ok, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || col >= D) return v;
  if (vec_ok) return ldg4(row + col);  // D % 4 == 0 -> whole group in range
  v.x = __ldg(row + col);
  if (col + 1 < D) v.y = __ldg(row + col + 1);
  if (col + 2 < D) v.z = __ldg(row + col + 2);
  if (col + 3 < D) v.w = __ldg(row + col + 3);
  return v;
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v, float4 s) {
  float4 c = make_float4(v.x - s.x, v.y - s.y, v.z - s.z, v.w - s.w);
  float4 h = make_float4(to_tf32(c.x), to_tf32(c.y), to_tf32(c.z), to_tf32(c.w));
  float4 l = make_float4(c.x - h.x, c.y - h.y, c.z - h.z, c.w - h.w);
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

// byte offset of (row r, 16-byte group g) inside an operand buffer laid out
// [g/8 chunk][r][128 B] with the 128B swizzle (16B unit index XOR (r & 7)).
__device__ __forceinline__ uint32_t op_offset(int g, int r) {
  return (uint32_t)((g >> 3) * CHUNK_BYTES + r * 128 + (((g & 7) ^ (r & 7)) << 4));
}

__device__ __forceinline__ void out_store(float* p, float v, int atomic_out) {
  if (atomic_out) atomicAdd(p, v); else *p = v;
}

__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_tf32x3_kernel(const GramParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_job;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const bool is_mma_warp = (warp == PROD_WARPS);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], PROD_WARPS);  // one arrive per producer warp
      mbar_init(&empty_bar[s], 1);          // one tcgen05.commit
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_fence_init();
  }
  if (is_mma_warp) tmem_alloc<TMEM_COLS>(&s_tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;

  uint32_t stage = 0, phase = 0;  // identical evolution in producers and the MMA warp
  uint32_t acc_phase = 0;

  const int D = P.D;
  const int jobs_total = P.C * P.T * P.KS;

  for (;;) {
    if (tid == 0) s_job = atomicAdd(P.job_counter, 1);
    __syncthreads();
    const int job = s_job;
    if (job >= jobs_total) break;

    // ---- decode job -> (class, tile, k-split) ----
    const int c = job / (P.T * P.KS);
    const int rem = job - c * (P.T * P.KS);
    int t = rem / P.KS;
    const int ks = rem - t * P.KS;
    int tm = 0;
    for (;; ++tm) {  // tiles of row-block tm: tn in [tm/2, TN)
      const int cnt = P.TN - (tm >> 1);
      if (t < cnt) break;
      t -= cnt;
    }
    const int tn = (tm >> 1) + t;
    const int m0 = tm * BM;
    const int n0 = tn * BN;
    int n_eff = D - n0;
    n_eff = n_eff > BN ? BN : ((n_eff + 15) & ~15);

    const int64_t row_begin = P.offsets[c];
    const int64_t n_c = P.offsets[c + 1] - row_begin;
    const int nkb_total = (int)((n_c + BK - 1) / BK);
    const int kb0 = (int)(((int64_t)ks * nkb_total) / P.KS);
    const int kb1 = (int)(((int64_t)(ks + 1) * nkb_total) / P.KS);
    const int nkb = kb1 - kb0;

    if (!is_mma_warp) {
      // =========================== producers ===========================
      // lane l owns float4 groups: A group l (cols m0+4l), B groups l and l+32.
      const int colA = m0 + 4 * lane;
      const int colB0 = n0 + 4 * lane;
      const int colB1 = n0 + 128 + 4 * lane;
      const bool vec = P.vec_ok != 0;
      float4 sA = make_float4(0.f, 0.f, 0.f, 0.f), sB0 = sA, sB1 = sA;
      if (P.shift != nullptr) {
        const float* sh = P.shift + (int64_t)c * D;
        sA = load_group(sh, colA, D, true, vec);
        sB0 = load_group(sh, colB0, D, true, vec);
        sB1 = load_group(sh, colB1, D, true, vec);
        // groups outside [0,D) load zeros and the data there is zero too -> contributes nothing
      }
      const int r0 = 2 * warp;  // this warp's two rows inside a stage

      float4 buf[2][6];
      auto issue = [&](int kb, float4(&b)[6]) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int64_t k = (int64_t)kb * BK + r0 + rr;
          const bool ok = k < n_c;
          const float* row = P.X;
          if (ok) row = P.X + (int64_t)__ldg(P.perm + row_begin + k) * P.ldx;
          b[rr * 3 + 0] = load_group(row, colA, D, ok, vec);
          b[rr * 3 + 1] = load_group(row, colB0, D, ok, vec);
          b[rr * 3 + 2] = load_group(row, colB1, D, ok, vec);
        }
      };
      auto consume = [&](int kb, float4(&b)[6]) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * STAGE_BYTES;
        uint8_t* a_hi = st;
        uint8_t* a_lo = st + A_BYTES;
        uint8_t* b_hi = st + 2 * A_BYTES;
        uint8_t* b_lo = st + 2 * A_BYTES + B_BYTES;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r0 + rr;
          const bool ok = ((int64_t)kb * BK + r) < n_c;
          // padded rows must be exactly zero (shift must not leak in)
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          split_store(a_hi, a_lo, op_offset(lane, r), b[rr * 3 + 0], ok ? sA : z);
          split_store(b_hi, b_lo, op_offset(lane, r), b[rr * 3 + 1], ok ? sB0 : z);
          split_store(b_hi, b_lo, op_offset(lane + 32, r), b[rr * 3 + 2], ok ? sB1 : z);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };

      if (nkb > 0) issue(kb0, buf[0]);
      if (nkb > 1) issue(kb0 + 1, buf[1]);
      for (int kb = kb0; kb < kb1; kb += 2) {
        consume(kb, buf[0]);
        if (kb + 2 < kb1) issue(kb + 2, buf[0]);
        if (kb + 1 < kb1) {
          consume(kb + 1, buf[1]);
          if (kb + 3 < kb1) issue(kb + 3, buf[1]);
        }
      }

      // =========================== epilogue ===========================
      const int q = warp & 3;   // TMEM lane quarter this warp may access
      const int h = warp >> 2;  // column half
      const int row = m0 + 32 * q + lane;
      float* grow = P.gram + ((int64_t)c * D + row) * D;
      if (nkb > 0) {
        mbar_wait(&tmem_full_bar, acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col0 = h * 128 + cc * 32;
          if (col0 >= n_eff) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
          tmem_ld_wait();
          if (row < D) {
            const int gc = n0 + col0;
            if (P.vec_ok && !P.atomic_out && gc + 32 <= D) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(grow + gc + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (gc + j < D) out_store(grow + gc + j, __uint_as_float(v[j]), P.atomic_out);
            }
          }
        }
        tc_fence_before_sync();
      } else if (!P.atomic_out) {
        // empty class (or empty split with plain stores): the tile is exactly zero
        if (row < D) {
          for (int cc = 0; cc < 4; ++cc) {
            const int gc = n0 + h * 128 + cc * 32;
            for (int j = 0; j < 32; ++j)
              if (gc + j < D) grow[gc + j] = 0.f;
          }
        }
      }
    } else {
      // =========================== MMA issuer ===========================
      const uint32_t idesc = make_idesc_tf32(BM, (uint32_t)n_eff, /*a MN-major*/ 1, /*b MN-major*/ 1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = st + 2 * A_BYTES + B_BYTES;
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            const uint32_t ko = k8 * 1024;  // next 8 samples = next swizzle atom in every chunk
            const uint64_t dA_hi = make_smem_desc(a_hi + ko, CHUNK_BYTES, 1024, LAYOUT_SW128);
            const uint64_t dA_lo = make_smem_desc(a_lo + ko, CHUNK_BYTES, 1024, LAYOUT_SW128);
            const uint64_t dB_hi = make_smem_desc(b_hi + ko, CHUNK_BYTES, 1024, LAYOUT_SW128);
            const uint64_t dB_lo = make_smem_desc(b_lo + ko, CHUNK_BYTES, 1024, LAYOUT_SW128);
            const uint32_t first = (kb > kb0 || k8 > 0) ? 1u : 0u;
            umma_tf32_ss(tmem_base, dA_lo, dB_hi, idesc, first);  // small terms first
            umma_tf32_ss(tmem_base, dA_hi, dB_lo, idesc, 1u);
            umma_tf32_ss(tmem_base, dA_hi, dB_hi, idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);  // frees the stage when these MMAs have read it
          if (kb == kb1 - 1) umma_commit(&tmem_full_bar);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    // accumulator drained, every stage consumed -> safe to start the next job
    __syncthreads();
  }

  tc_fence_before_sync();
  __syncthreads();
  if (is_mma_warp) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// UMMA probe: one CTA, D[128 x N] = A^T B for A [K x 128], B [K x N] row-major fp32 in global,
// staged into shared memory with a caller-chosen canonical layout / descriptor. Used by
// tests/ and tools/ to pin the operand layout assumptions of gram_tf32x3_kernel on hardware.
//   mode 0: MN-major, 128B swizzle, [chunk][k][128B]      (what the Gram kernel uses)
//   mode 1: K-major, no swizzle, core matrices 8(mn) x 16B, [k/4][mn/8][8][16B]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const float* A, const float* B, float* Dout, int K, int N, int mode, uint32_t lbo, uint32_t sbo,
                  uint32_t layout_type, uint32_t a_major, uint32_t b_major, uint32_t kstep_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sA = smem;                       // up to 128 cols x K
  uint8_t* sB = smem + 128 * K * 4;         // N cols x K
  // stage operands
  for (int idx = tid; idx < K * 128; idx += blockDim.x) {
    const int k = idx / 128, m = idx % 128;
    uint32_t off;
    if (mode == 0) off = (m / 32) * (K * 128) + k * 128 + ((((m % 32) / 4) ^ (k & 7)) << 4) + (m % 4) * 4;
    else off = (k / 4) * (128 * 16) + (m / 8) * 128 + (m % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float*>(sA + off) = A[idx];
  }
  for (int idx = tid; idx < K * N; idx += blockDim.x) {
    const int k = idx / N, n = idx % N;
    uint32_t off;
    if (mode == 0) off = (n / 32) * (K * 128) + k * 128 + ((((n % 32) / 4) ^ (k & 7)) << 4) + (n % 4) * 4;
    else off = (k / 4) * (N * 16) + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float*>(sB + off) = B[idx];
  }
  if (tid == 0) { mbar_init(&done_bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(&s_tmem_base);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, (uint32_t)N, a_major, b_major);
      for (int k8 = 0; k8 < K / 8; ++k8) {
        const uint64_t dA = make_smem_desc(smem_u32(sA) + k8 * kstep_bytes, lbo, sbo, layout_type);
        const uint64_t dB = make_smem_desc(smem_u32(sB) + k8 * (mode == 0 ? kstep_bytes : (kstep_bytes / 128) * N),
                                           mode == 0 ? lbo : (lbo / 128) * N, sbo, layout_type);
        umma_tf32_ss(tmem_base, dA, dB, idesc, k8 > 0 ? 1u : 0u);
      }
      umma_commit(&done_bar);
    }
    __syncwarp();
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after_sync();
  for (int col0 = 0; col0 < N; col0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)col0, v);
    tmem_ld_wait();
    const int row = 32 * warp + lane;
    for (int j = 0; j < 32; ++j)
      if (col0 + j < N) Dout[row * N + col0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

}  // namespace

int gram_tiles_per_class(int D, int* TM_out, int* TN_out) {
  const int TM = (D + BM - 1) / BM, TN = (D + BN - 1) / BN;
  int T = 0;
  for (int tm = 0; tm < TM; ++tm) T += TN - (tm >> 1);
  if (TM_out) *TM_out = TM;
  if (TN_out) *TN_out = TN;
  return T;
}

cudaError_t launch_class_gram(const float* X, int64_t ldx, const int32_t* perm, const int64_t* offsets,
                              const float* shift, int D, int C, float* gram, int accumulate, int ksplit,
                              int* job_counter, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gram_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  GramParams P;
  P.X = X; P.ldx = ldx; P.perm = perm; P.offsets = offsets; P.shift = shift; P.gram = gram;
  P.job_counter = job_counter; P.D = D; P.C = C;
  P.T = gram_tiles_per_class(D, &P.TM, &P.TN);
  P.KS = ksplit < 1 ? 1 : ksplit;
  P.atomic_out = (accumulate || P.KS > 1) ? 1 : 0;
  P.vec_ok = (D % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
             ((reinterpret_cast<uintptr_t>(gram) & 15) == 0) &&
             (shift == nullptr || (reinterpret_cast<uintptr_t>(shift) & 15) == 0);
  cudaError_t e = cudaMemsetAsync(job_counter, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  const int64_t jobs = (int64_t)C * P.T * P.KS;
  if (jobs <= 0) return cudaSuccess;
  const int grid = (int)(jobs < num_sms ? jobs : num_sms);
  gram_tf32x3_kernel<<<grid, GRAM_THREADS, GRAM_SMEM, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_umma_probe(const float* A, const float* B, float* Dout, int K, int N, int mode, uint32_t lbo,
                              uint32_t sbo, uint32_t layout_type, uint32_t a_major, uint32_t b_major,
                              uint32_t kstep_bytes, cudaStream_t stream) {
  const int smem = (128 + N) * K * 4 + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_kernel<<<1, 128, smem, stream>>>(A, B, Dout, K, N, mode, lbo, sbo, layout_type, a_major, b_major,
                                              kstep_bytes);
  return cudaGetLastError();
}

}  // namespace sqfa
