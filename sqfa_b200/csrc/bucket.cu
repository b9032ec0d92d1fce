// K1: label bucketing. Stable LSB radix sort (8-bit digits) of (label, row id) on device.
//
// Replaces the C passes of `(labels == i).nonzero().squeeze(1)` in the reference
// (/root/reference/src/sqfa/statistics.py:36-37): for every class i the reference takes the row
// ids with that label in ascending order -- i.e. the stable bucket permutation. The result here is
// bit-exact equal to torch.sort(labels, stable=True).indices restricted to labels in [0, C);
// rows whose label is outside [0, C) belong to no class in the reference and are parked in a
// trailing "dropped" bucket with key C.
#include <cstdint>
#include <atomic>
#include <cuda_runtime.h>

#include "sqfa_internal.h"

namespace sqfa {

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int WARPS_PER_BLOCK = 8;
constexpr int STEPS_PER_TILE = 32;
constexpr int TILE = 32 * STEPS_PER_TILE;  // elements per warp tile

__device__ __forceinline__ int32_t clip_label(int64_t y, int32_t C) { return (y < 0 || y >= C) ? C : (int32_t)y; }

// max(labels) without touching a copy engine: block maxima meet in a scratch slot (device globals,
// one of LM_SLOTS per call so that calls on different streams do not share one), the last block
// publishes the result to `out` -- device memory or MAPPED PINNED HOST memory (then the caller
// waits on an event instead of issuing a device-to-host copy that would queue behind bulk
// transfers of other streams) -- and resets the slot for its next use.
constexpr int LM_SLOTS = 64;
__device__ long long g_lm_max[LM_SLOTS] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
__device__ unsigned int g_lm_count[LM_SLOTS];  // zero-initialised

__global__ void label_max_kernel(const int64_t* __restrict__ labels, int64_t n, int slot, long long* out) {
  __shared__ long long s_m[8];
  long long m = -1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long v = labels[i];
    m = v > m ? v : m;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const long long other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other > m ? other : m;
  }
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = s_m[w] > m ? s_m[w] : m;
    atomicMax(&g_lm_max[slot], m);
    __threadfence();
    if (atomicAdd(&g_lm_count[slot], 1u) == gridDim.x - 1) {  // last block: publish and reset the slot
      __threadfence();
      const long long r = (long long)atomicExch(reinterpret_cast<unsigned long long*>(&g_lm_max[slot]), ~0ull);
      g_lm_count[slot] = 0u;
      *reinterpret_cast<volatile long long*>(out) = r;
      __threadfence_system();
    }
  }
}

template <bool FIRST>
__device__ __forceinline__ int32_t load_key(const int64_t* labels, const int32_t* keys, int64_t i, int32_t C) {
  if (FIRST) return clip_label(labels[i], C);
  return keys[i];
}

template <bool FIRST>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
radix_hist_kernel(const int64_t* __restrict__ labels, const int32_t* __restrict__ keys, int64_t n, int32_t C,
                  int shift, int32_t* __restrict__ hist, int ntiles) {
  __shared__ int32_t s_hist[WARPS_PER_BLOCK][RADIX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * WARPS_PER_BLOCK + warp;
  if (tile >= ntiles) return;
  int32_t* h = s_hist[warp];
  for (int b = lane; b < RADIX; b += 32) h[b] = 0;
  __syncwarp();
  const int64_t base = (int64_t)tile * TILE;
  for (int s = 0; s < STEPS_PER_TILE; ++s) {
    const int64_t i = base + s * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = valid ? ((uint32_t)load_key<FIRST>(labels, keys, i, C) >> shift) & (RADIX - 1) : RADIX;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && (__ffs(peers) - 1) == lane) h[d] += __popc(peers);
    __syncwarp();
  }
  for (int b = lane; b < RADIX; b += 32) hist[(int64_t)b * ntiles + tile] = h[b];
}

// Exclusive scan of the (digit-major) histogram in three coalesced kernels: per-chunk sums, scan of
// the chunk sums (one block), per-chunk scan with the chunk offset. A chunk is 8192 entries: every
// thread of a 1024-thread block owns 8 consecutive ones (two int4 loads).
constexpr int SCAN_THREADS = 1024, SCAN_PER_THREAD = 8, SCAN_CHUNK = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ void scan_load8(const int32_t* a, int64_t n, int64_t i0, int32_t (&v)[8]) {
  if (i0 + 8 <= n) {
    const int4 x = *reinterpret_cast<const int4*>(a + i0), y = *reinterpret_cast<const int4*>(a + i0 + 4);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = i0 + j < n ? a[i0 + j] : 0;
  }
}

// block-wide exclusive scan of one value per thread (1024 threads); returns the block total in *total
__device__ __forceinline__ int32_t block_exclusive_scan(int32_t x, int32_t* s_warp, int32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc = x;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;           // exclusive offset of warp `lane`
    if (lane == 31) s_warp[32] = wi;  // block total
  }
  __syncthreads();
  const int32_t r = s_warp[warp] + inc - x;
  *total = s_warp[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_chunk_sums_kernel(const int32_t* __restrict__ a, int64_t n,
                                                                       int32_t* __restrict__ sums) {
  __shared__ int32_t s_warp[33];
  int32_t v[8];
  scan_load8(a, n, (int64_t)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD, v);
  int32_t t = 0, total;
#pragma unroll
  for (int j = 0; j < 8; ++j) t += v[j];
  block_exclusive_scan(t, s_warp, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int32_t* sums, int nchunks) {
  __shared__ int32_t s_warp[33];
  int32_t carry = 0;
  for (int base = 0; base < nchunks; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const int32_t x = i < nchunks ? sums[i] : 0;
    int32_t total;
    const int32_t e = block_exclusive_scan(x, s_warp, &total);
    if (i < nchunks) sums[i] = carry + e;
    carry += total;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int32_t* a, int64_t n,
                                                                  const int32_t* __restrict__ sums) {
  __shared__ int32_t s_warp[33];
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
  int32_t v[8];
  scan_load8(a, n, i0, v);
  int32_t t = 0, total;
#pragma unroll
  for (int j = 0; j < 8; ++j) t += v[j];
  int32_t run = sums[blockIdx.x] + block_exclusive_scan(t, s_warp, &total);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int32_t x = v[j];
    v[j] = run;
    run += x;
  }
  if (i0 + 8 <= n) {
    *reinterpret_cast<int4*>(a + i0) = make_int4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<int4*>(a + i0 + 4) = make_int4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (i0 + j < n) a[i0 + j] = v[j];
  }
}

static void launch_exclusive_scan(int32_t* a, int64_t n, int32_t* chunk_sums, cudaStream_t stream) {
  const int nchunks = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
  scan_chunk_sums_kernel<<<nchunks, SCAN_THREADS, 0, stream>>>(a, n, chunk_sums);
  scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(chunk_sums, nchunks);
  scan_apply_kernel<<<nchunks, SCAN_THREADS, 0, stream>>>(a, n, chunk_sums);
}

template <bool FIRST>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
radix_scatter_kernel(const int64_t* __restrict__ labels, const int32_t* __restrict__ keys_in,
                     const int32_t* __restrict__ vals_in, int64_t n, int32_t C, int shift,
                     const int32_t* __restrict__ hist_scanned, int ntiles, int32_t* __restrict__ keys_out,
                     int32_t* __restrict__ vals_out) {
  __shared__ int32_t s_base[WARPS_PER_BLOCK][RADIX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * WARPS_PER_BLOCK + warp;
  if (tile >= ntiles) return;
  int32_t* bp = s_base[warp];
  for (int b = lane; b < RADIX; b += 32) bp[b] = hist_scanned[(int64_t)b * ntiles + tile];
  __syncwarp();
  const int64_t base = (int64_t)tile * TILE;
  const uint32_t lt = (1u << lane) - 1u;
  for (int s = 0; s < STEPS_PER_TILE; ++s) {
    const int64_t i = base + s * 32 + lane;
    const bool valid = i < n;
    int32_t key = 0, val = 0;
    uint32_t d = RADIX;
    if (valid) {
      key = load_key<FIRST>(labels, keys_in, i, C);
      val = FIRST ? (int32_t)i : vals_in[i];
      d = ((uint32_t)key >> shift) & (RADIX - 1);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    int32_t pos = 0;
    if (valid) pos = bp[d] + __popc(peers & lt);  // stable: earlier lanes first
    __syncwarp();
    if (valid && (__ffs(peers) - 1) == lane) bp[d] += __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
  }
}

// offsets[c] = first sorted position whose key >= c, for c in [0, C+1]; offsets[C+1] = n.
__global__ void class_offsets_kernel(const int32_t* __restrict__ sorted_keys, int64_t n, int32_t C,
                                     int64_t* __restrict__ offsets, int64_t* __restrict__ counts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  const int32_t prev = (i == 0) ? -1 : sorted_keys[i - 1];
  const int32_t cur = (i == n) ? C + 1 : sorted_keys[i];
  for (int32_t c = prev + 1; c <= cur; ++c) offsets[c] = i;
}

__global__ void class_counts_kernel(const int64_t* __restrict__ offsets, int32_t C, int64_t* __restrict__ counts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= C) counts[c] = offsets[c + 1] - offsets[c];
}

// One 8-bit pass (at most 255 classes) with a histogram of at most MID_MAX_ENTRIES entries: the exclusive scan
// of the digit-major histogram in ONE block (coalesced chunks of 8192 entries with a running carry), which also writes the
// class offsets and counts -- class c is digit c, its first row sits at the scanned entry of (digit c, tile 0).
// Replaces the three scan launches and the two offset launches of the general path.
constexpr int MID_MAX_ENTRIES = 1 << 16;  // 256 tiles = 262 144 rows (8 chunks: beyond that the parallel scan wins)

__global__ void __launch_bounds__(SCAN_THREADS) scan_offsets_kernel(int32_t* __restrict__ hist, int total, int ntiles,
                                                                     int32_t C, int64_t n,
                                                                     int64_t* __restrict__ counts,
                                                                     int64_t* __restrict__ offsets) {
  __shared__ int32_t s_warp[33];
  const int tid = threadIdx.x;
  int32_t carry = 0;
  for (int base = 0; base < total; base += SCAN_CHUNK) {  // chunks of 8192 entries, 8 consecutive ones per thread
    const int64_t i0 = (int64_t)base + tid * SCAN_PER_THREAD;
    int32_t v[8];
    scan_load8(hist, total, i0, v);
    int32_t t = 0, chunk_total;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += v[j];
    int32_t run = carry + block_exclusive_scan(t, s_warp, &chunk_total);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int32_t x = v[j];
      v[j] = run;
      run += x;
    }
    if (i0 + 8 <= total) {
      *reinterpret_cast<int4*>(hist + i0) = make_int4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<int4*>(hist + i0 + 4) = make_int4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (i0 + j < total) hist[i0 + j] = v[j];
    }
    carry += chunk_total;
  }
  __syncthreads();  // (block-wide visibility of the global writes above)
  for (int c = tid; c <= C + 1; c += SCAN_THREADS) offsets[c] = (c <= C) ? (int64_t)hist[(int64_t)c * ntiles] : n;
  for (int c = tid; c <= C; c += SCAN_THREADS) {
    const int64_t nxt = (c + 1 <= C) ? (int64_t)hist[(int64_t)(c + 1) * ntiles] : n;
    counts[c] = nxt - (int64_t)hist[(int64_t)c * ntiles];
  }
}

// Small problems (at most SMALL_MAX_TILES tiles, at most 255 classes: one 8-bit pass) in ONE block: per-tile
// histograms, their scan, offsets / counts and the stable scatter all stay in shared memory -- launch
// latency outweighs the work at this size. Beyond about 8 000 rows the single block is the slower choice
// (c2, 50 000 rows: 59 us against three launches of the one-pass path above).
constexpr int SMALL_MAX_TILES = 8;

__global__ void __launch_bounds__(1024) bucket_small_kernel(const int64_t* __restrict__ labels, int n, int32_t C,
                                                            int ntiles, int64_t* __restrict__ counts,
                                                            int64_t* __restrict__ offsets, int32_t* __restrict__ perm) {
  extern __shared__ int32_t s_hist[];  // [RADIX][ntiles], digit-major like the general path
  __shared__ int32_t s_warp[33];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = RADIX * ntiles;
  for (int i = tid; i < total; i += 1024) s_hist[i] = 0;
  __syncthreads();
  for (int tile = warp; tile < ntiles; tile += 32) {  // a tile's column is private to its warp
    const int base = tile * TILE;
    for (int s = 0; s < STEPS_PER_TILE; ++s) {
      const int i = base + s * 32 + lane;
      const bool valid = i < n;
      const uint32_t d = valid ? (uint32_t)clip_label(labels[i], C) : RADIX;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      if (valid && (__ffs(peers) - 1) == lane) s_hist[d * ntiles + tile] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  // exclusive scan over the digit-major histogram: a contiguous run of entries per thread
  const int per = (total + 1023) / 1024;
  const int lo = tid * per, hi = min(total, lo + per);
  int32_t sum = 0;
  for (int i = lo; i < hi; ++i) sum += s_hist[i];
  int32_t tot;
  int32_t run = block_exclusive_scan(sum, s_warp, &tot);
  for (int i = lo; i < hi; ++i) {
    const int32_t t = s_hist[i];
    s_hist[i] = run;
    run += t;
  }
  __syncthreads();
  // class c is digit c: its first row sits at the scanned entry of (digit c, tile 0)
  for (int c = tid; c <= C + 1; c += 1024) offsets[c] = (c <= C) ? (int64_t)s_hist[c * ntiles] : (int64_t)n;
  for (int c = tid; c <= C; c += 1024) {
    const int64_t nxt = (c + 1 <= C) ? (int64_t)s_hist[(c + 1) * ntiles] : (int64_t)n;
    counts[c] = nxt - (int64_t)s_hist[c * ntiles];
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
  for (int tile = warp; tile < ntiles; tile += 32) {
    const int base = tile * TILE;
    for (int s = 0; s < STEPS_PER_TILE; ++s) {
      const int i = base + s * 32 + lane;
      const bool valid = i < n;
      const uint32_t d = valid ? (uint32_t)clip_label(labels[i], C) : RADIX;
      const uint32_t peers = __match_any_sync(0xffffffffu, d);
      int32_t pos = 0;
      if (valid) pos = s_hist[d * ntiles + tile] + __popc(peers & lt);  // stable: earlier lanes first
      __syncwarp();
      if (valid && (__ffs(peers) - 1) == lane) s_hist[d * ntiles + tile] += __popc(peers);
      __syncwarp();
      if (valid) perm[pos] = i;
    }
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

cudaError_t launch_label_max(const int64_t* labels, int64_t n, int64_t* out_max, cudaStream_t stream) {
  static std::atomic<unsigned> next_slot{0};
  const int slot = (int)(next_slot.fetch_add(1u) % LM_SLOTS);
  const int threads = 256;
  int64_t blocks = (n + threads * 8 - 1) / (threads * 8);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;  // n == 0: one block publishes -1
  label_max_kernel<<<(int)blocks, threads, 0, stream>>>(labels, n, slot, reinterpret_cast<long long*>(out_max));
  return cudaGetLastError();
}

size_t bucket_workspace_bytes(int64_t n, int32_t C) {
  (void)C;
  const int64_t ntiles = (n + TILE - 1) / TILE;
  size_t b = 0;
  b += 4 * align_up((size_t)(n > 0 ? n : 1) * sizeof(int32_t), 256);  // keys x2, vals x2
  const size_t hist_entries = (size_t)(ntiles > 0 ? ntiles : 1) * RADIX;
  b += align_up(hist_entries * sizeof(int32_t), 256);
  b += align_up(((hist_entries + SCAN_CHUNK - 1) / SCAN_CHUNK) * sizeof(int32_t), 256);  // scan chunk sums
  return b;
}

cudaError_t launch_bucket_labels(const int64_t* labels, int64_t n, int32_t C, int64_t* counts, int64_t* offsets,
                                 int32_t* perm, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (ws_bytes < bucket_workspace_bytes(n, C)) return cudaErrorInvalidValue;
  if (n >= (int64_t)1 << 31) return cudaErrorInvalidValue;
  const int ntiles = (int)((n + TILE - 1) / TILE);
  if (n > 0 && ntiles <= SMALL_MAX_TILES && C + 1 <= RADIX) {
    const int smem = RADIX * ntiles * (int)sizeof(int32_t);
    static int smem_set[kMaxDevices] = {0};
    cudaError_t e = ensure_dynamic_smem(bucket_small_kernel, smem, smem_set);
    if (e != cudaSuccess) return e;
    bucket_small_kernel<<<1, 1024, smem, stream>>>(labels, (int)n, C, ntiles, counts, offsets, perm);
    return cudaGetLastError();
  }
  uint8_t* p = static_cast<uint8_t*>(ws);
  const size_t nb = align_up((size_t)(n > 0 ? n : 1) * sizeof(int32_t), 256);
  int32_t* keys[2] = {reinterpret_cast<int32_t*>(p), reinterpret_cast<int32_t*>(p + nb)};
  int32_t* vals[2] = {reinterpret_cast<int32_t*>(p + 2 * nb), reinterpret_cast<int32_t*>(p + 3 * nb)};
  int32_t* hist = reinterpret_cast<int32_t*>(p + 4 * nb);
  int32_t* chunk_sums = reinterpret_cast<int32_t*>(
      p + 4 * nb + align_up((size_t)(ntiles > 0 ? ntiles : 1) * RADIX * sizeof(int32_t), 256));

  if (n > 0 && C + 1 <= RADIX && (int64_t)ntiles * RADIX <= MID_MAX_ENTRIES) {
    // one 8-bit pass: histogram, single-block scan (+ offsets, counts), stable scatter straight into perm
    const int blocks = (ntiles + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int threads = WARPS_PER_BLOCK * 32;
    radix_hist_kernel<true><<<blocks, threads, 0, stream>>>(labels, nullptr, n, C, 0, hist, ntiles);
    scan_offsets_kernel<<<1, SCAN_THREADS, 0, stream>>>(hist, ntiles * RADIX, ntiles, C, n, counts, offsets);
    radix_scatter_kernel<true><<<blocks, threads, 0, stream>>>(labels, nullptr, nullptr, n, C, 0, hist, ntiles,
                                                               keys[1], perm);
    return cudaGetLastError();
  }

  int bits = 0;
  while ((1ll << bits) < (int64_t)C + 1) ++bits;  // keys take values 0..C
  int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
  if (passes < 1) passes = 1;

  const int32_t* sorted_keys = nullptr;
  if (n > 0) {
    const int blocks = (ntiles + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    const int threads = WARPS_PER_BLOCK * 32;
    int cur = 0;
    for (int pass = 0; pass < passes; ++pass) {
      const int shift = pass * RADIX_BITS;
      // the last pass writes the row ids straight into the caller's perm buffer
      int32_t* vout = (pass == passes - 1) ? perm : vals[cur ^ 1];
      if (pass == 0) {
        radix_hist_kernel<true><<<blocks, threads, 0, stream>>>(labels, nullptr, n, C, shift, hist, ntiles);
        launch_exclusive_scan(hist, (int64_t)ntiles * RADIX, chunk_sums, stream);
        radix_scatter_kernel<true><<<blocks, threads, 0, stream>>>(labels, nullptr, nullptr, n, C, shift, hist,
                                                                   ntiles, keys[cur ^ 1], vout);
      } else {
        radix_hist_kernel<false><<<blocks, threads, 0, stream>>>(nullptr, keys[cur], n, C, shift, hist, ntiles);
        launch_exclusive_scan(hist, (int64_t)ntiles * RADIX, chunk_sums, stream);
        radix_scatter_kernel<false><<<blocks, threads, 0, stream>>>(nullptr, keys[cur], vals[cur], n, C, shift,
                                                                    hist, ntiles, keys[cur ^ 1], vout);
      }
      cur ^= 1;
    }
    sorted_keys = keys[cur];
  }
  {
    const int threads = 256;
    const int64_t blocks = (n + 1 + threads - 1) / threads;
    class_offsets_kernel<<<(int)blocks, threads, 0, stream>>>(sorted_keys, n, C, offsets, counts);
    class_counts_kernel<<<(C + 1 + threads - 1) / threads, threads, 0, stream>>>(offsets, C, counts);
  }
  return cudaGetLastError();
}

}  // namespace sqfa
