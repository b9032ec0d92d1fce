// K4 on the tensor cores for 8 < k <= 32 filters: T_c = F S_c as the skinny GEMM  T_c^T = S_c F^T
// (S_c is symmetric), tcgen05.mma kind::tf32 with the 3xTF32 split, accumulators in TMEM.
//
// Reference path replaced: conjugate_matrix(S, F) (/root/reference/src/sqfa/linalg.py:41) through
// transform_scatters (model.py:187) -- its first factor, the only part of a closure evaluation that reads
// C D^2 floats. The SIMT kernel (project.cu, project_stream) does k FMAs per loaded float: HBM-bound for
// k <= 8, but FP32-FMA bound beyond (c4, k = 16: 3.8 TB/s; k = 32: 1.6 TB/s -- the SIMT pipes cannot reach the
// HBM roofline at k = 32 at all). Here a CTA streams a tile of 128 rows of S_c through shared memory:
//   A = the rows (M = 128, K = columns j of S, K-major exactly as they lie in memory),
//   B = the filters (N = k, K-major exactly as F lies in memory),
// both split x = hi + lo (hi = TF32-rounded) by the producer warps on their way from registers to shared
// memory, and per K = 8 step TWO MMAs give the three products of the split:
//   A_hi x [B_hi ; B_lo]  (N = 2 KP: hi.hi and hi.lo side by side)      A_lo x B_hi  (N = KP).
// The three partial accumulators are added by the epilogue warps (fp32), chain by chain of 512 columns
// (the tensor core truncates when it accumulates: short chains keep the result at fp32 level, DESIGN.md
// section 4), and the tile of T is stored transposed back into (k, D) layout -- 128-byte rows per filter.
// The MMA work is a fraction of the time the tile's bytes take to arrive from HBM (tensor pipe 7 % active).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int PT_ROWS = 128;                       // rows of S per tile (MMA M, TMEM lanes)
constexpr int PT_BK = 32;                          // columns of S per stage (128 bytes per row)
constexpr int PT_STAGES = 4;                      // two super-stages of two stages
constexpr int PT_CHAIN = 16;                       // stages per accumulation chain (512 columns)
constexpr int PT_PROD_WARPS = 8, PT_EPI_WARPS = 4;
constexpr int PT_MMA_WARP = PT_PROD_WARPS;
constexpr int PT_THREADS = (PT_PROD_WARPS + 1 + PT_EPI_WARPS) * 32;
// K-major, no swizzle (gram.cu / tools/umma_probe.py): element (row r, column chunk q of 4) of an operand at
// q * LBO + (r / 8) * 128 + (r % 8) * 16. For A the chunks are 16 bytes further apart than they need to be
// (LBO = 2048 + 16): a quarter warp that stores the 8 chunks of ONE row then hits 8 different 16-byte bank
// groups instead of one (the producers read whole 512-byte row segments, see below).
constexpr uint32_t PT_A_LBO = PT_ROWS * 16 + 16, PT_SBO = 128;
constexpr int PT_A_BYTES = (PT_BK / 4) * PT_A_LBO;  // 16.1 KB: one of A_hi / A_lo
constexpr int PT_SUPER = 2;                          // stages a producer loads at once: 256 contiguous bytes per row

template <int KP>
struct PtGeom {
  static constexpr int B_ROWS = 2 * KP;                        // [B_hi ; B_lo]
  static constexpr int B_BYTES = B_ROWS * PT_BK * 4;
  static constexpr uint32_t B_LBO = B_ROWS * 16;
  static constexpr int STAGE_BYTES = ((2 * PT_A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int SMEM = PT_STAGES * STAGE_BYTES + 1024;
  static constexpr uint32_t ACC_COLS = 3 * KP;                  // hi.hi | hi.lo | lo.hi
  static constexpr uint32_t TMEM_COLS = 2 * ACC_COLS <= 64 ? 64 : (2 * ACC_COLS <= 128 ? 128 : 256);
};

// byte offset of element (row r of the operand, K index kk) in a K-major operand of `lbo` bytes per 4-element chunk
__device__ __forceinline__ uint32_t op_offset(int r, int kchunk, uint32_t lbo) {
  return (uint32_t)kchunk * lbo + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u;
}

template <int KP>
__global__ void __launch_bounds__(PT_THREADS, 1)
project_tc_kernel(const float* __restrict__ S, const float* __restrict__ F, int C, int D, int k, float* __restrict__ T) {
  using G = PtGeom<KP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[PT_STAGES];   // producers -> MMA (8 warp arrivals)
  __shared__ __align__(8) uint64_t empty_bar[PT_STAGES];  // MMA (commit) -> producers
  __shared__ __align__(8) uint64_t acc_full_bar[2];       // MMA (commit) -> epilogue: a chain is complete
  __shared__ __align__(8) uint64_t acc_empty_bar[2];      // epilogue (4 warp arrivals) -> MMA
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < PT_STAGES; ++s) {
      mbar_init(&full_bar[s], PT_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full_bar[b], 1);
      mbar_init(&acc_empty_bar[b], PT_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == PT_MMA_WARP) tmem_alloc<G::TMEM_COLS>(&s_tmem_base);
  // filter rows beyond k (KP is k rounded up to 16) are exact zeros in every stage, written once
  for (int i = tid; i < PT_STAGES * G::B_BYTES / 16; i += PT_THREADS) {
    const int stage = i / (G::B_BYTES / 16), q = i % (G::B_BYTES / 16);
    st_shared_v4(smem_u32(smem + stage * G::STAGE_BYTES + 2 * PT_A_BYTES) + 16u * q, make_float4(0.f, 0.f, 0.f, 0.f));
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;

  const int row_tiles = (D + PT_ROWS - 1) / PT_ROWS;
  const int njobs = C * row_tiles;
  // per job; the columns are padded with zeros to a multiple of 64 (whole super-stages)
  const int nstages = PT_SUPER * ((D + PT_SUPER * PT_BK - 1) / (PT_SUPER * PT_BK));
  const int nchains = (nstages + PT_CHAIN - 1) / PT_CHAIN;

  if (warp < PT_PROD_WARPS) {
    // =========================== producers ===========================
    // A super-stage is 128 rows x 64 columns = two stages of the ring. A warp instruction loads 256 contiguous
    // bytes of each of two adjacent rows (lane l: row l / 16 of the pair, columns 4 (l % 16) ..): rows are
    // D * 4 bytes apart in DRAM, and 64-byte pieces per row and instruction ran at 2.3 TB/s (no DRAM page
    // locality). Lane l's 16 bytes are chunk l % 8 of stage (l / 8) % 2: with the padded LBO a quarter warp
    // stores 8 chunks of one row without a bank conflict.
    // B (the filters): thread t < 8 KP loads chunk (t / 8) % 8 of filter t % 8 + 8 (t / 64), once per stage.
    // Register pipeline (the scoreboard rule of gram.cu: the first use of ANY loaded register waits for ALL
    // loads in flight): two full register sets; the first unpack of a set is the one wait, the loads of the next
    // super-stage go out right behind it into the other set and fly while this one is split and stored.
    // Loads are unconditional with clamped addresses; what lies outside the matrix is zeroed when consumed.
    const int nsuper = nstages / PT_SUPER;
    const int my_jobs = (njobs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_jobs * nsuper;  // super-stages of this CTA
    const bool has_b = tid < KP * 8;
    const int bchunk4 = 4 * ((tid >> 3) & 7), brow = (tid & 7) + 8 * (tid >> 6);
    const bool b_row_ok = has_b && brow < k;
    const float* bptr = F + (size_t)(b_row_ok ? brow : 0) * D;
    const uint32_t bh_off = op_offset(brow, bchunk4 >> 2, G::B_LBO), bl_off = op_offset(KP + brow, bchunk4 >> 2, G::B_LBO);
    const int lane_row = lane >> 4, lane_col4 = 4 * (lane & 15);
    const uint32_t a_lane_off = (uint32_t)((lane >> 3) & 1) * G::STAGE_BYTES + (uint32_t)(lane & 7) * PT_A_LBO;
    const uint32_t sbase = smem_u32(smem);
    struct Set { b128_t a[8], b[PT_SUPER]; };
    struct Cursor { int job, ss, r0; const float* Sc; };
    auto start = [&](Cursor& cu) {
      cu.job = (int)blockIdx.x; cu.ss = 0;
      const int c = cu.job / row_tiles;
      cu.r0 = (cu.job - c * row_tiles) * PT_ROWS;
      cu.Sc = S + (size_t)c * D * D;
    };
    auto advance = [&](Cursor& cu) {
      if (++cu.ss == nsuper) {
        cu.ss = 0;
        cu.job += (int)gridDim.x;
        if (cu.job < njobs) {
          const int c = cu.job / row_tiles;
          cu.r0 = (cu.job - c * row_tiles) * PT_ROWS;
          cu.Sc = S + (size_t)c * D * D;
        }
      }
    };
    auto issue = [&](Set& R, const Cursor& cu) {
      const int j = min(cu.ss * PT_SUPER * PT_BK + lane_col4, D - 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = min(cu.r0 + 2 * (warp + 8 * i) + lane_row, D - 1);
        R.a[i] = ldg_nc_b128(cu.Sc + (size_t)r * D + j);
      }
#pragma unroll
      for (int sg = 0; sg < PT_SUPER; ++sg)
        R.b[sg] = ldg_nc_b128(bptr + min((cu.ss * PT_SUPER + sg) * PT_BK + bchunk4, D - 4));
    };
    auto split = [](float4 v, float4& h, float4& l) {
      h.x = to_tf32(v.x); l.x = v.x - h.x;
      h.y = to_tf32(v.y); l.y = v.y - h.y;
      h.z = to_tf32(v.z); l.z = v.z - h.z;
      h.w = to_tf32(v.w); l.w = v.w - h.w;
    };
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    Cursor lc, cc;  // load cursor (one super-stage ahead) and consume cursor
    Set P, Q;
    uint32_t slot = 0, phase = 0;  // ring slot (stages 2 slot, 2 slot + 1) and its barrier phase
    auto consume = [&](Set& R, Set& Next, bool more) {
      float4 first = unpack_b128(R.a[0]);  // THE wait: every load of this set has landed behind it
      if (more) {
        issue(Next, lc);
        advance(lc);
      }
      mbar_wait(&empty_bar[2 * slot], phase ^ 1);
      mbar_wait(&empty_bar[2 * slot + 1], phase ^ 1);
      const bool jok = cc.ss * PT_SUPER * PT_BK + lane_col4 < D;  // D % 4 == 0: a chunk is inside or outside
      const uint32_t so = sbase + (2 * slot) * G::STAGE_BYTES + a_lane_off;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = 2 * (warp + 8 * i) + lane_row;
        float4 v = i == 0 ? first : unpack_b128(R.a[i]);
        if (!(jok && cc.r0 + row < D)) v = z;
        float4 h, l;
        split(v, h, l);
        const uint32_t off = so + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
        st_shared_v4(off, h);
        st_shared_v4(off + PT_A_BYTES, l);
      }
      if (has_b) {
#pragma unroll
        for (int sg = 0; sg < PT_SUPER; ++sg) {
          float4 v = unpack_b128(R.b[sg]);
          if (!(b_row_ok && (cc.ss * PT_SUPER + sg) * PT_BK + bchunk4 < D)) v = z;
          float4 h, l;
          split(v, h, l);
          const uint32_t sb = sbase + (2 * slot + sg) * G::STAGE_BYTES + 2 * PT_A_BYTES;
          st_shared_v4(sb + bh_off, h);
          st_shared_v4(sb + bl_off, l);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full_bar[2 * slot]);
        mbar_arrive(&full_bar[2 * slot + 1]);
      }
      advance(cc);
      if (++slot == 2) { slot = 0; phase ^= 1; }
    };
    if (total > 0) {
      start(lc);
      start(cc);
      issue(P, lc);
      advance(lc);
    }
    for (int t = 0; t < total; t += 2) {
      consume(P, Q, t + 1 < total);
      if (t + 1 < total) consume(Q, P, t + 2 < total);
    }
  } else if (warp == PT_MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc_wide = make_idesc_tf32(PT_ROWS, 2 * KP, 0, 0), idesc_narrow = make_idesc_tf32(PT_ROWS, KP, 0, 0);
    uint32_t stage = 0, phase = 0, chain = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      for (int ch = 0; ch < nchains; ++ch, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait(&acc_empty_bar[buf], ((chain >> 1) & 1u) ^ 1u);  // the epilogue has read this buffer's last chain
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + buf * G::ACC_COLS;
        const int s_end = min(nstages, (ch + 1) * PT_CHAIN);
        for (int st = ch * PT_CHAIN; st < s_end; ++st) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * G::STAGE_BYTES);
            const uint32_t a_hi = sa, a_lo = sa + PT_A_BYTES, b = sa + 2 * PT_A_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < PT_BK / 8; ++k8) {  // one MMA consumes two 4-column chunks
              const uint64_t dA_hi = make_smem_desc(a_hi + k8 * 2 * PT_A_LBO, PT_A_LBO, PT_SBO, 0);
              const uint64_t dA_lo = make_smem_desc(a_lo + k8 * 2 * PT_A_LBO, PT_A_LBO, PT_SBO, 0);
              const uint64_t dB = make_smem_desc(b + k8 * 2 * G::B_LBO, G::B_LBO, PT_SBO, 0);
              const uint32_t accumulate = (st > ch * PT_CHAIN || k8 > 0) ? 1u : 0u;
              umma_tf32_ss(acc, dA_hi, dB, idesc_wide, accumulate);             // hi.hi | hi.lo
              umma_tf32_ss(acc + 2 * KP, dA_lo, dB, idesc_narrow, accumulate);  // lo.hi (the first KP rows of B)
            }
            umma_commit(&empty_bar[stage]);  // the stage may be refilled once these MMAs have read it
            if (st == s_end - 1) umma_commit(&acc_full_bar[buf]);
          }
          __syncwarp();
          if (++stage == PT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t chain = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int c = job / row_tiles, r0 = (job - c * row_tiles) * PT_ROWS;
      float run[KP];
#pragma unroll
      for (int f = 0; f < KP; ++f) run[f] = 0.f;
      for (int ch = 0; ch < nchains; ++ch, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait(&acc_full_bar[buf], (chain >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t acc = tq + buf * G::ACC_COLS;
#pragma unroll
        for (int f0 = 0; f0 < KP; f0 += 16) {
          uint32_t hh[16], hl[16], lh[16];
          tmem_ld_32x32b_x16(acc + f0, hh);
          tmem_ld_32x32b_x16(acc + KP + f0, hl);
          tmem_ld_32x32b_x16(acc + 2 * KP + f0, lh);
          tmem_ld_wait();
#pragma unroll
          for (int f = 0; f < 16; ++f)
            run[f0 + f] += __uint_as_float(hh[f]) + (__uint_as_float(hl[f]) + __uint_as_float(lh[f]));
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[buf]);
      }
      // T[c][f][r0 + 32 q + lane]: a warp writes 128 contiguous bytes per filter
      const int r = r0 + 32 * q + lane;
      if (r < D) {
        float* out = T + (size_t)c * k * D + r;
#pragma unroll
        for (int f = 0; f < KP; ++f)
          if (f < k) out[(size_t)f * D] = run[f];
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == PT_MMA_WARP) tmem_dealloc<G::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// Variant with the raw tile staged by TMA: ONE cp.async.bulk.tensor.2d per stage brings the 128 x 32 box of
// S_c (and one the KP x 32 box of F) into a raw ring of four stages -- 66 KB per SM in flight without holding
// a register -- and reports the bytes to an mbarrier; the eight splitter warps read the raw rows back
// (a quarter warp reads the 8 chunks of one row: 128 contiguous bytes), split them and store hi / lo into a
// two-stage operand ring (conflict-free thanks to the padded LBO); MMA issuer and epilogue as above. Rows
// and columns beyond the matrix are zero-filled by the copy unit or zeroed by the splitters. Costs one more
// write and one more read of every byte in shared memory (the trade DESIGN.md section 4 describes for the
// Gram). MEASURED (B200, SQFA_PROJECT_TC_BULK=1): c4 (k = 16) 391 us, c5 shape (k = 32) 210 us against 297 /
// 155 us for the register-staged producers above and 274 / 261 us for the SIMT pass -- the second pass through
// shared memory and the two-stage operand ring cost more than the deeper prefetch gains, so this variant is
// opt-in; it is what DESIGN.md section 4 argues for the Gram, here as a measured A/B. (A first version with
// one 1-D bulk copy per ROW, 160 copies of 128 bytes per stage, took 2.0 ms at c4: the copy unit needs ~70
// cycles per copy -- small bulk copies are not a substitute for a tensor map.)
// ------------------------------------------------------------------------------------------------
constexpr int PB_RAW_STAGES = 4, PB_OP_STAGES = 2;
constexpr int PB_RAW_ROW = PT_BK * 4;                // dense 128-byte rows, as the copy unit writes a box
constexpr int PB_THREADS = PT_THREADS + 32;          // + the copy warp
constexpr int PB_COPY_WARP = PT_PROD_WARPS + 1 + PT_EPI_WARPS;

template <int KP>
struct PbGeom {
  static constexpr int RAW_A_BYTES = PT_ROWS * PB_RAW_ROW, RAW_B_BYTES = KP * PB_RAW_ROW;
  static constexpr int RAW_BYTES = RAW_A_BYTES + RAW_B_BYTES;  // rows of the tile, then the filters
  static constexpr int SMEM = PB_OP_STAGES * PtGeom<KP>::STAGE_BYTES + PB_RAW_STAGES * RAW_BYTES + 1024;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

template <int KP>
__global__ void __launch_bounds__(PB_THREADS, 1)
project_tcb_kernel(const __grid_constant__ CUtensorMap map_S, const __grid_constant__ CUtensorMap map_F, int C, int D,
                   int k, float* __restrict__ T) {
  using G = PtGeom<KP>;
  using GB = PbGeom<KP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* raw = smem + PB_OP_STAGES * G::STAGE_BYTES;
  __shared__ __align__(8) uint64_t raw_full_bar[PB_RAW_STAGES];   // copy warp (expect_tx) + the copies' bytes
  __shared__ __align__(8) uint64_t raw_empty_bar[PB_RAW_STAGES];  // splitters (8 warp arrivals) -> copy warp
  __shared__ __align__(8) uint64_t full_bar[PB_OP_STAGES];        // splitters -> MMA
  __shared__ __align__(8) uint64_t empty_bar[PB_OP_STAGES];       // MMA (commit) -> splitters
  __shared__ __align__(8) uint64_t acc_full_bar[2];
  __shared__ __align__(8) uint64_t acc_empty_bar[2];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < PB_RAW_STAGES; ++s) {
      mbar_init(&raw_full_bar[s], 1);
      mbar_init(&raw_empty_bar[s], PT_PROD_WARPS);
    }
    for (int s = 0; s < PB_OP_STAGES; ++s) {
      mbar_init(&full_bar[s], PT_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full_bar[b], 1);
      mbar_init(&acc_empty_bar[b], PT_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == PT_MMA_WARP) tmem_alloc<G::TMEM_COLS>(&s_tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = s_tmem_base;

  const int row_tiles = (D + PT_ROWS - 1) / PT_ROWS;
  const int njobs = C * row_tiles;
  const int nstages = (D + PT_BK - 1) / PT_BK;  // per job
  const int nchains = (nstages + PT_CHAIN - 1) / PT_CHAIN;

  if (warp == PB_COPY_WARP) {
    // =========================== TMA issuer (one thread) ===========================
    if (lane == 0) {
      uint32_t rs = 0, rphase = 0;
      for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        const int c = job / row_tiles, r0 = (job - c * row_tiles) * PT_ROWS;
        for (int st = 0; st < nstages; ++st) {
          mbar_wait(&raw_empty_bar[rs], rphase ^ 1);
          mbar_arrive_expect_tx(&raw_full_bar[rs], (uint32_t)GB::RAW_BYTES);  // a box counts in full, zero fill included
          const uint32_t dst = smem_u32(raw + rs * GB::RAW_BYTES);
          tma_load_2d(dst, &map_S, st * PT_BK, c * D + r0, &raw_full_bar[rs]);
          tma_load_2d(dst + GB::RAW_A_BYTES, &map_F, st * PT_BK, 0, &raw_full_bar[rs]);
          if (++rs == PB_RAW_STAGES) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp < PT_PROD_WARPS) {
    // =========================== splitters ===========================
    // lane l: chunk l % 8 of row l / 8 of a group of four rows; warp w takes the row groups w, w + 8, w + 16,
    // w + 24. A quarter warp reads 128 contiguous bytes and writes the 8 chunks of one row (padded LBO).
    const int chunk = lane & 7, row_lane = lane >> 3;
    const bool has_b = tid < KP * 8;
    const int brow = tid >> 3;  // tid < 8 KP: filter row, chunk tid % 8 = lane % 8
    const bool b_row_ok = has_b && brow < k;
    const uint32_t bh_off = op_offset(brow, chunk, G::B_LBO), bl_off = op_offset(KP + brow, chunk, G::B_LBO);
    auto split = [](float4 v, float4& h, float4& l) {
      h.x = to_tf32(v.x); l.x = v.x - h.x;
      h.y = to_tf32(v.y); l.y = v.y - h.y;
      h.z = to_tf32(v.z); l.z = v.z - h.z;
      h.w = to_tf32(v.w); l.w = v.w - h.w;
    };
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t rs = 0, rphase = 0, stage = 0, phase = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int r0 = (job % row_tiles) * PT_ROWS;
      for (int st = 0; st < nstages; ++st) {
        mbar_wait(&raw_full_bar[rs], rphase);
        const uint32_t src = smem_u32(raw + rs * GB::RAW_BYTES) + 16u * chunk;
        float4 va[4], vb = z;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = 4 * (warp + 8 * u) + row_lane;
          va[u] = ld_shared_v4(src + row * PB_RAW_ROW);
          if (r0 + row >= D) va[u] = z;  // rows of the next class (or zero fill behind the last one)
        }
        if (has_b) {
          vb = ld_shared_v4(src + GB::RAW_A_BYTES + brow * PB_RAW_ROW);
          if (!b_row_ok) vb = z;
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        const uint32_t sa = smem_u32(smem + stage * G::STAGE_BYTES);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = 4 * (warp + 8 * u) + row_lane;
          float4 h, l;
          split(va[u], h, l);
          const uint32_t off = sa + op_offset(row, chunk, PT_A_LBO);
          st_shared_v4(off, h);
          st_shared_v4(off + PT_A_BYTES, l);
        }
        if (has_b) {
          float4 h, l;
          split(vb, h, l);
          st_shared_v4(sa + 2 * PT_A_BYTES + bh_off, h);
          st_shared_v4(sa + 2 * PT_A_BYTES + bl_off, l);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&full_bar[stage]);
          mbar_arrive(&raw_empty_bar[rs]);  // every value of the raw stage has been consumed
        }
        if (++stage == PB_OP_STAGES) { stage = 0; phase ^= 1; }
        if (++rs == PB_RAW_STAGES) { rs = 0; rphase ^= 1; }
      }
    }
  } else if (warp == PT_MMA_WARP) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc_wide = make_idesc_tf32(PT_ROWS, 2 * KP, 0, 0), idesc_narrow = make_idesc_tf32(PT_ROWS, KP, 0, 0);
    uint32_t stage = 0, phase = 0, chain = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      for (int ch = 0; ch < nchains; ++ch, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait(&acc_empty_bar[buf], ((chain >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + buf * G::ACC_COLS;
        const int s_end = min(nstages, (ch + 1) * PT_CHAIN);
        for (int st = ch * PT_CHAIN; st < s_end; ++st) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after_sync();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * G::STAGE_BYTES);
            const uint32_t a_hi = sa, a_lo = sa + PT_A_BYTES, b = sa + 2 * PT_A_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < PT_BK / 8; ++k8) {
              const uint64_t dA_hi = make_smem_desc(a_hi + k8 * 2 * PT_A_LBO, PT_A_LBO, PT_SBO, 0);
              const uint64_t dA_lo = make_smem_desc(a_lo + k8 * 2 * PT_A_LBO, PT_A_LBO, PT_SBO, 0);
              const uint64_t dB = make_smem_desc(b + k8 * 2 * G::B_LBO, G::B_LBO, PT_SBO, 0);
              const uint32_t accumulate = (st > ch * PT_CHAIN || k8 > 0) ? 1u : 0u;
              umma_tf32_ss(acc, dA_hi, dB, idesc_wide, accumulate);
              umma_tf32_ss(acc + 2 * KP, dA_lo, dB, idesc_narrow, accumulate);
            }
            umma_commit(&empty_bar[stage]);
            if (st == s_end - 1) umma_commit(&acc_full_bar[buf]);
          }
          __syncwarp();
          if (++stage == PB_OP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;
    const uint32_t tq = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t chain = 0;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
      const int c = job / row_tiles, r0 = (job - c * row_tiles) * PT_ROWS;
      float run[KP];
#pragma unroll
      for (int f = 0; f < KP; ++f) run[f] = 0.f;
      for (int ch = 0; ch < nchains; ++ch, ++chain) {
        const uint32_t buf = chain & 1u;
        mbar_wait(&acc_full_bar[buf], (chain >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t acc = tq + buf * G::ACC_COLS;
#pragma unroll
        for (int f0 = 0; f0 < KP; f0 += 16) {
          uint32_t hh[16], hl[16], lh[16];
          tmem_ld_32x32b_x16(acc + f0, hh);
          tmem_ld_32x32b_x16(acc + KP + f0, hl);
          tmem_ld_32x32b_x16(acc + 2 * KP + f0, lh);
          tmem_ld_wait();
#pragma unroll
          for (int f = 0; f < 16; ++f)
            run[f0 + f] += __uint_as_float(hh[f]) + (__uint_as_float(hl[f]) + __uint_as_float(lh[f]));
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty_bar[buf]);
      }
      const int r = r0 + 32 * q + lane;
      if (r < D) {
        float* out = T + (size_t)c * k * D + r;
#pragma unroll
        for (int f = 0; f < KP; ++f)
          if (f < k) out[(size_t)f * D] = run[f];
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == PT_MMA_WARP) tmem_dealloc<G::TMEM_COLS>(tmem_base);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link dependency on libcuda):
// a 2-D float32 tensor of `rows` x `cols` (row-major), boxes of box_rows x box_cols, no swizzle, zero fill
static cudaError_t encode_map_2d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                                 uint32_t box_rows) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    fn = reinterpret_cast<encode_fn>(p);
  }
  const cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int KP>
cudaError_t run_project_tcb(const float* S, const float* F, int C, int D, int k, float* T, cudaStream_t st) {
  using GB = PbGeom<KP>;
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(project_tcb_kernel<KP>, GB::SMEM, smem_set);
    if (e != cudaSuccess) return e;
  }
  CUtensorMap map_S, map_F;
  cudaError_t e = encode_map_2d(&map_S, S, (uint64_t)D, (uint64_t)C * D, PT_BK, PT_ROWS);
  if (e != cudaSuccess) return e;
  if ((e = encode_map_2d(&map_F, F, (uint64_t)D, (uint64_t)k, PT_BK, KP)) != cudaSuccess) return e;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int njobs = C * ((D + PT_ROWS - 1) / PT_ROWS);
  const int grid = njobs < sms ? njobs : sms;
  project_tcb_kernel<KP><<<grid, PB_THREADS, GB::SMEM, st>>>(map_S, map_F, C, D, k, T);
  return cudaGetLastError();
}

template <int KP>
cudaError_t run_project_tc(const float* S, const float* F, int C, int D, int k, float* T, cudaStream_t st) {
  using G = PtGeom<KP>;
  static int smem_set[kMaxDevices] = {0};
  {
    cudaError_t e = ensure_dynamic_smem(project_tc_kernel<KP>, G::SMEM, smem_set);
    if (e != cudaSuccess) return e;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int njobs = C * ((D + PT_ROWS - 1) / PT_ROWS);
  const int grid = njobs < sms ? njobs : sms;
  project_tc_kernel<KP><<<grid, PT_THREADS, G::SMEM, st>>>(S, F, C, D, k, T);
  return cudaGetLastError();
}

}  // namespace

// T[c] = F S[c] for all classes on the tensor cores; valid for 8 < k <= 32, D % 4 == 0, 16-byte aligned S and F.
// Taken for k > 16, where it wins as built (c5 shape, k = 32: 155 us against 261 us for the SIMT pass, 2.7 TB/s);
// at k = 16 the SIMT pass is still ahead (c4: 274 against 297 us): what bounds this kernel is the bytes its
// producers keep in flight (one 33 KB super-stage per SM in registers: ncu stall_long_sb 63 %), not the tensor
// pipe (7 % active). SQFA_PROJECT_TC=0 / 1 forces it off / on for every 8 < k <= 32.
bool project_tc_applicable(const float* S, const float* F, int D, int k) {
  const char* env = getenv("SQFA_PROJECT_TC");  // read per call: tests switch it
  const int mode = env == nullptr ? -1 : atoi(env);
  if (mode == 0 || k <= 8 || k > 32 || D % 4 != 0 || D < PT_BK ||
      ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(F)) & 15) != 0)
    return false;
  return mode == 1 || k > 16;
}

cudaError_t launch_project_tc(const float* S, const float* F, int C, int D, int k, float* T, cudaStream_t st) {
  if (C <= 0) return cudaSuccess;
  const char* env = getenv("SQFA_PROJECT_TC_BULK");  // 1: raw tile staged by TMA (A/B switch)
  if (env != nullptr && atoi(env) != 0) {
    if (k <= 16) return run_project_tcb<16>(S, F, C, D, k, T, st);
    return run_project_tcb<32>(S, F, C, D, k, T, st);
  }
  if (k <= 16) return run_project_tc<16>(S, F, C, D, k, T, st);
  return run_project_tc<32>(S, F, C, D, k, T, st);
}

}  // namespace sqfa
