// TEST-ONLY object (built into tests/native/libsqfa_probe.so, not part of libsqfa_b200.so or its header).
// Pins the tcgen05 operand-layout assumptions of the Gram kernel on hardware: one CTA,
// D[128 x N] = A^T B through one tcgen05.mma chain with a caller-chosen shared-memory layout /
// descriptor. Used by tests/test_hp1_gpu.py and tools/umma_probe.py.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ptx.cuh"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

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

extern "C" int sqfa_debug_umma_probe(const float* A, const float* B, float* Dout, int32_t K, int32_t N, int32_t mode,
                                     uint32_t lbo, uint32_t sbo, uint32_t layout_type, uint32_t a_major,
                                     uint32_t b_major, uint32_t kstep_bytes, void* stream) {
  if (A == nullptr || B == nullptr || Dout == nullptr || K <= 0 || K % 8 != 0 || K > 64 || N < 16 || N > 256 ||
      N % 32 != 0)
    return -1;
  return (int)sqfa::launch_umma_probe(A, B, Dout, K, N, mode, lbo, sbo, layout_type, a_major, b_major, kstep_bytes,
                                      static_cast<cudaStream_t>(stream));
}
