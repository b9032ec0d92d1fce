// K2: segmented per-class Gram accumulation  G_c = sum_{i in class c} (x_i - s_c)(x_i - s_c)^T
// on the sm_100a tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM) as 3xTF32
// split precision:  x = hi + lo,  G ~= hi^T hi + hi^T lo + lo^T hi   (fp32-level accuracy).
//
// Replaces the per-class loop of the reference (gather `points[indices]`, centre, einsum
// "ij,ik->jk"):  /root/reference/src/sqfa/statistics.py:36-47 and :113-122.
//
// Data flow per CTA (persistent, one CTA per SM, dynamic job counter):
//   12 producer warps: one thread per operand column gathers the class rows through the bucket
//                      permutation (coalesced 128 B per warp and row) straight into registers
//                      (no raw staging in smem -- shared-memory bandwidth is the binding resource
//                      for 3xTF32), subtracts the class shift, splits hi/lo and
//                      stores both as 16-byte chunks of the K-major UMMA operand layout,
//                      fence.proxy.async, arrive on full[stage].
//   1 MMA warp       : one elected lane issues 3 tcgen05.mma per K=8 step (lo*hi, hi*lo, hi*hi),
//                      tcgen05.commit -> empty[stage]; after the last K block commit -> tmem_full.
//   epilogue         : the producer warps read the 128 x N accumulator with tcgen05.ld and store
//                      (or red.add when the class is split along K / accumulating) to gram.
//
// Only tiles that intersect the upper triangle are computed; K3 mirrors them.
//
// Accuracy note (measured on B200, tools/exp_gram.py): the tensor core truncates when it adds
// into the fp32 accumulator, a bias of about -2^-25 of the accumulator per tcgen05.mma. Two
// counter-measures keep the Gram at fp32 level: (1) the tiny cross terms (lo*hi, hi*lo) go to their
// OWN accumulator (TMEM columns 256..511) so only one accumulate per K=8 step hits the large
// one; (2) a class is cut into chains of at most CHAIN_ROWS samples (device-side job plan), every
// chain accumulates from zero and is added to `gram` with fp32 round-to-nearest red.global.add.
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
constexpr int PROD_WARPS = (BM + BN) / 32;  // 12: one producer thread per operand column
constexpr int GRAM_THREADS = (PROD_WARPS + 1) * 32;

// Operand layout in shared memory: K-major, no swizzle ("interleaved" core matrices), pinned on
// hardware by tools/umma_probe.py:  element (column c, sample k) of an operand of width W lives at
//   (k/4) * (W*16)  +  (c/8) * 128  +  (c%8) * 16  +  (k%4) * 4      bytes,
// i.e. 8x(4 samples) core matrices of 128 contiguous bytes; LBO = W*16 (next 4 samples),
// SBO = 128 (next 8 columns). One tcgen05.mma (K = 8 tf32) consumes two k-chunks.
constexpr int A_BYTES = BM * BK * 4;                    //  8 KB
constexpr int B_BYTES = BN * BK * 4;                    // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // hi+lo for A and B = 48 KB
constexpr int GRAM_SMEM = STAGES * STAGE_BYTES + 1024;  // + alignment slack
constexpr uint32_t TMEM_COLS = 512;   // [0,256): hi*hi accumulator, [256,512): cross-term accumulator
constexpr uint32_t TMEM_SMALL = 256;
constexpr int CHAIN_ROWS = 512;        // samples per accumulation chain (64 accumulates -> bias < 2e-6)
constexpr uint32_t A_LBO = BM * 16, B_LBO = BN * 16, OP_SBO = 128;
constexpr uint32_t LAYOUT_NONE = 0;

__device__ float g_zero[4] = {0.f, 0.f, 0.f, 0.f};  // what out-of-range operand columns read

struct GramParams {
  const float* X;
  int64_t ldx;
  const int32_t* perm;       // rows sorted by class (stable)
  const int64_t* offsets;    // C+1 class offsets into perm
  const float* shift;        // C x D (or nullptr -> 0)
  float* gram;               // C x D x D
  int* job_counter;          // [0] dynamic job counter
  const int* job_base;       // C+1 prefix sums of jobs per class (built by gram_plan_kernel)
  int D;
  int C;
  int TM, TN, T;             // tile grid and tiles per class
  int chain_rows;            // samples per accumulation chain
  int vec_ok;                // gram rows 16-byte aligned -> red.global.add.v4.f32 in the epilogue
};

// jobs of class c = chains_c * T, chains_c = max(1, ceil(n_c / chain_rows));  job_base = exclusive scan
__global__ void __launch_bounds__(1024) gram_plan_kernel(const int64_t* __restrict__ offsets, int C, int T,
                                                         int chain_rows, int* __restrict__ job_base) {
  __shared__ int s_sum[1024];
  const int tid = threadIdx.x;
  const int per = (C + 1023) / 1024;
  const int lo = tid * per, hi = min(C, lo + per);
  int sum = 0;
  for (int c = lo; c < hi; ++c) {
    const int64_t n_c = offsets[c + 1] - offsets[c];
    const int chains = n_c > 0 ? (int)((n_c + chain_rows - 1) / chain_rows) : 1;
    sum += chains * T;
  }
  s_sum[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = tid >= o ? s_sum[tid - o] : 0;
    __syncthreads();
    s_sum[tid] += v;
    __syncthreads();
  }
  int run = s_sum[tid] - sum;
  for (int c = lo; c < hi; ++c) {
    const int64_t n_c = offsets[c + 1] - offsets[c];
    const int chains = n_c > 0 ? (int)((n_c + chain_rows - 1) / chain_rows) : 1;
    job_base[c] = run;
    run += chains * T;
  }
  if (tid == 1023) job_base[C] = s_sum[1023];
}

__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_tf32x3_kernel(const GramParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t s_tmem_base;
  __shared__ int s_job[4];  // class, tile, chain, chains of the class  (class < 0: no more work)

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

  for (;;) {
    if (tid == 0) {
      // ---- fetch and decode job -> (class, chain, tile); tiles of one chain are adjacent jobs so
      //      concurrently running CTAs share the gathered rows in L2 ----
      const int job = atomicAdd(P.job_counter, 1);
      if (job >= __ldg(P.job_base + P.C)) {
        s_job[0] = -1;
      } else {
        int lo = 0, hi = P.C;  // largest c with job_base[c] <= job
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(P.job_base + mid) <= job) lo = mid; else hi = mid;
        }
        const int rem = job - __ldg(P.job_base + lo);
        s_job[0] = lo;
        s_job[1] = rem % P.T;
        s_job[2] = rem / P.T;
        s_job[3] = (__ldg(P.job_base + lo + 1) - __ldg(P.job_base + lo)) / P.T;
      }
    }
    __syncthreads();
    const int c = s_job[0];
    if (c < 0) break;
    int t = s_job[1];
    const int chain = s_job[2], chains = s_job[3];
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
    const int kb0 = (int)(((int64_t)chain * nkb_total) / chains);
    const int kb1 = (int)(((int64_t)(chain + 1) * nkb_total) / chains);
    const int nkb = kb1 - kb0;

    if (!is_mma_warp) {
      // =========================== producers ===========================
      // thread <-> one operand column: warps 0-3 the 128 A columns, warps 4-11 the 256 B columns.
      // Per stage it gathers that column of the 16 sample rows (coalesced 128 B per warp and row),
      // centres, splits hi/lo and writes 4+4 16-byte chunks (4 consecutive samples each).
      const bool isA = warp < BM / 32;
      const int cw = isA ? (warp * 32 + lane) : (warp * 32 + lane - BM);  // column inside the operand
      const int col = (isA ? m0 : n0) + cw;                               // column of X
      const bool col_ok = col < D && (isA || cw < n_eff);
      // The inner loops are instruction-issue bound, so everything per-thread is folded into three
      // values: a column base pointer, a row stride in bytes and the centring shift. Columns
      // outside the matrix read a device zero with stride 0 and shift 0 -> exact zeros, no selects.
      const char* xcol = col_ok ? reinterpret_cast<const char*>(P.X + col) : reinterpret_cast<const char*>(g_zero);
      const uint32_t ldb = col_ok ? (uint32_t)(P.ldx * 4) : 0u;
      const float sh = (P.shift != nullptr && col_ok) ? __ldg(P.shift + (int64_t)c * D + col) : 0.f;
      const uint32_t op_lbo = isA ? A_LBO : B_LBO;
      uint8_t* const hi_base = smem + (isA ? 0u : 2u * A_BYTES) + (uint32_t)((cw >> 3) * 128 + (cw & 7) * 16);
      uint8_t* const lo_base = hi_base + (isA ? A_BYTES : B_BYTES);
      const int32_t* const permc = P.perm + row_begin;
      const int lane16 = lane & (BK - 1);

      float buf[3][BK];  // three stages of loads in flight per thread (registers)
      // lane r < 16 holds the row id of sample r of the NEXT stage to issue (fetched a stage ahead);
      // samples past the end of the chain's class read row 0 and are masked in consume_tail().
      auto load_row = [&](int kb) -> uint32_t {
        const int64_t k = (int64_t)kb * BK + lane16;
        return (kb < kb1 && k < n_c) ? (uint32_t)__ldg(permc + k) : 0u;
      };
      uint32_t nextrow = load_row(kb0);
      auto issue = [&](int kb, float(&b)[BK]) {
        const uint32_t myrow = nextrow;
        nextrow = load_row(kb + 1);
#pragma unroll
        for (int r = 0; r < BK; ++r) {  // SHFL + IMAD.WIDE.U32 + LDG per element, nothing else
          const uint32_t row = __shfl_sync(0xffffffffu, myrow, r);
          uint64_t addr;  // one IMAD.WIDE.U32: base + row * stride
          asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(addr) : "r"(row), "r"(ldb), "l"(xcol));
          b[r] = __ldg(reinterpret_cast<const float*>(addr));
        }
      };
      auto store_stage = [&](const float(&x)[BK]) {
        uint8_t* hp = hi_base + stage * STAGE_BYTES;
        uint8_t* lp = lo_base + stage * STAGE_BYTES;
#pragma unroll
        for (int kc = 0; kc < BK / 4; ++kc) {
          float4 h, l;
          h.x = to_tf32(x[4 * kc + 0]); l.x = x[4 * kc + 0] - h.x;
          h.y = to_tf32(x[4 * kc + 1]); l.y = x[4 * kc + 1] - h.y;
          h.z = to_tf32(x[4 * kc + 2]); l.z = x[4 * kc + 2] - h.z;
          h.w = to_tf32(x[4 * kc + 3]); l.w = x[4 * kc + 3] - h.w;
          *reinterpret_cast<float4*>(hp + kc * op_lbo) = h;
          *reinterpret_cast<float4*>(lp + kc * op_lbo) = l;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      };
      auto consume = [&](int kb, float(&b)[BK]) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const int64_t nv = n_c - (int64_t)kb * BK;  // valid rows of this stage
        float x[BK];
        if (nv >= BK) {  // full stage (all but the last of a class): FADD, CVT, FADD per element
#pragma unroll
          for (int r = 0; r < BK; ++r) x[r] = b[r] - sh;
        } else {
#pragma unroll
          for (int r = 0; r < BK; ++r) x[r] = (r < (int)nv) ? b[r] - sh : 0.f;
        }
        store_stage(x);
      };

      if (nkb > 0) issue(kb0, buf[0]);
      if (nkb > 1) issue(kb0 + 1, buf[1]);
      if (nkb > 2) issue(kb0 + 2, buf[2]);
      for (int kb = kb0; kb < kb1; kb += 3) {
        consume(kb, buf[0]);
        if (kb + 3 < kb1) issue(kb + 3, buf[0]);
        if (kb + 1 < kb1) {
          consume(kb + 1, buf[1]);
          if (kb + 4 < kb1) issue(kb + 4, buf[1]);
        }
        if (kb + 2 < kb1) {
          consume(kb + 2, buf[2]);
          if (kb + 5 < kb1) issue(kb + 5, buf[2]);
        }
      }

      // =========================== epilogue ===========================
      // warp w may read TMEM lanes [32*(w%4), +32); the three warps of a lane quarter take the
      // 32-column chunks cc = w/4, w/4 + 3, w/4 + 6.
      const int q = warp & 3;
      const int row = m0 + 32 * q + lane;
      float* grow = P.gram + ((int64_t)c * D + row) * D;
      if (nkb > 0) {
        mbar_wait(&tmem_full_bar, acc_phase);
        acc_phase ^= 1;
        tc_fence_after_sync();
#pragma unroll 1
        for (int cc = warp >> 2; cc < BN / 32; cc += PROD_WARPS / 4) {
          const int col0 = cc * 32;
          if (col0 >= n_eff) break;  // warp-uniform
          uint32_t v[32], w[32];
          const uint32_t ta = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)col0;
          tmem_ld_32x32b_x32(ta, v);
          tmem_ld_32x32b_x32(ta + TMEM_SMALL, w);
          tmem_ld_wait();
          if (row < D) {
            const int gc = n0 + col0;
            if (P.vec_ok && gc + 32 <= D) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                atomicAdd(reinterpret_cast<float4*>(grow + gc + j),
                          make_float4(__uint_as_float(v[j]) + __uint_as_float(w[j]),
                                      __uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1]),
                                      __uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2]),
                                      __uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3])));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (gc + j < D) atomicAdd(grow + gc + j, __uint_as_float(v[j]) + __uint_as_float(w[j]));
            }
          }
        }
        tc_fence_before_sync();
      }
    } else {
      // =========================== MMA issuer ===========================
      const uint32_t idesc = make_idesc_tf32(BM, (uint32_t)n_eff, /*a K-major*/ 0, /*b K-major*/ 0);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = st + 2 * A_BYTES + B_BYTES;
#pragma unroll
          for (int k8 = 0; k8 < BK / 8; ++k8) {
            const uint32_t ka = k8 * 2 * A_LBO, kbo = k8 * 2 * B_LBO;  // 8 samples = two k-chunks
            const uint64_t dA_hi = make_smem_desc(a_hi + ka, A_LBO, OP_SBO, LAYOUT_NONE);
            const uint64_t dA_lo = make_smem_desc(a_lo + ka, A_LBO, OP_SBO, LAYOUT_NONE);
            const uint64_t dB_hi = make_smem_desc(b_hi + kbo, B_LBO, OP_SBO, LAYOUT_NONE);
            const uint64_t dB_lo = make_smem_desc(b_lo + kbo, B_LBO, OP_SBO, LAYOUT_NONE);
            const uint32_t acc = (kb > kb0 || k8 > 0) ? 1u : 0u;
            umma_tf32_ss(tmem_base + TMEM_SMALL, dA_lo, dB_hi, idesc, acc);  // cross terms: own accumulator
            umma_tf32_ss(tmem_base + TMEM_SMALL, dA_hi, dB_lo, idesc, 1u);
            umma_tf32_ss(tmem_base, dA_hi, dB_hi, idesc, acc);
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
//   mode 2: DECODE A: A's smem is filled with its own word index (mod 2048, exact in tf32) and
//           read through the caller's descriptor; B is a K-major selector B[k][n] = (n == k), so
//           Dout[m][n<8] = word index the hardware fetched for A(k = n, m). K must be 8.
//   mode 3: DECODE B: the same with the roles swapped: Dout[m<8][n] = word index of B(k = m, n).
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
  if (mode >= 2) {
    // raw-filled operand: 8192 words; selector operand after it
    float* raw = reinterpret_cast<float*>(smem);
    for (int w = tid; w < 8192; w += blockDim.x) raw[w] = (float)(w & 2047);
    float* sel = raw + 8192;  // K-major no-swizzle selector of width W: [k/4][W/8][8][4]
    const int W = (mode == 2) ? N : 128;
    for (int idx = tid; idx < 8 * W; idx += blockDim.x) {
      const int k = idx / W, x = idx % W;
      sel[(k / 4) * (W * 4) + (x / 8) * 32 + (x % 8) * 4 + (k % 4)] = (x == k) ? 1.f : 0.f;
    }
    if (tid == 0) { mbar_init(&done_bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<256>(&s_tmem_base);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tb = s_tmem_base;
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t dRaw = make_smem_desc(smem_u32(raw), lbo, sbo, layout_type);
        const uint64_t dSel = make_smem_desc(smem_u32(sel), (uint32_t)W * 16, 128, 0);
        if (mode == 2)
          umma_tf32_ss(tb, dRaw, dSel, make_idesc_tf32(128, (uint32_t)N, a_major, 0), 0u);
        else
          umma_tf32_ss(tb, dSel, dRaw, make_idesc_tf32(128, (uint32_t)N, 0, b_major), 0u);
        umma_commit(&done_bar);
      }
      __syncwarp();
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after_sync();
    for (int col0 = 0; col0 < N; col0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tb + ((uint32_t)(32 * warp) << 16) + (uint32_t)col0, v);
      tmem_ld_wait();
      const int row = 32 * warp + lane;
      for (int j = 0; j < 32; ++j)
        if (col0 + j < N) Dout[row * N + col0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tb);
    return;
  }
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
                              const float* shift, int D, int C, float* gram, int accumulate, int chain_rows,
                              int* ws, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gram_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GRAM_SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (C <= 0) return cudaSuccess;
  GramParams P;
  P.X = X; P.ldx = ldx; P.perm = perm; P.offsets = offsets; P.shift = shift; P.gram = gram;
  P.job_counter = ws; P.job_base = ws + 4; P.D = D; P.C = C;
  P.T = gram_tiles_per_class(D, &P.TM, &P.TN);
  P.chain_rows = chain_rows > 0 ? ((chain_rows + BK - 1) / BK) * BK : CHAIN_ROWS;
  P.vec_ok = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(gram) & 15) == 0);
  cudaError_t e = cudaMemsetAsync(ws, 0, 4 * sizeof(int), stream);
  if (e != cudaSuccess) return e;
  if (!accumulate) {  // chains are summed into gram with red.add -> start from zero
    e = cudaMemsetAsync(gram, 0, (size_t)C * D * D * sizeof(float), stream);
    if (e != cudaSuccess) return e;
  }
  gram_plan_kernel<<<1, 1024, 0, stream>>>(offsets, C, P.T, P.chain_rows, ws + 4);
  gram_tf32x3_kernel<<<num_sms, GRAM_THREADS, GRAM_SMEM, stream>>>(P);
  return cudaGetLastError();
}

size_t gram_workspace_bytes(int C) { return (size_t)(C + 1 + 4) * sizeof(int); }

cudaError_t launch_umma_probe(const float* A, const float* B, float* Dout, int K, int N, int mode, uint32_t lbo,
                              uint32_t sbo, uint32_t layout_type, uint32_t a_major, uint32_t b_major,
                              uint32_t kstep_bytes, cudaStream_t stream) {
  const int smem = (mode >= 2 ? 8192 * 4 + 8 * 256 * 4 : (128 + N) * K * 4) + 1024;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_kernel<<<1, 128, smem, stream>>>(A, B, Dout, K, N, mode, lbo, sbo, layout_type, a_major, b_major,
                                              kstep_bytes);
  return cudaGetLastError();
}

}  // namespace sqfa
