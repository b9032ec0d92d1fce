// The whole closure body of the SQFA fitting loop as ONE native call:
//   loss = -mean_{i>j} d(E_i, E_j),  E_c = embed(F S_c F^T + noise I [, F m_c]),  dF = dloss/dF
// (reference: /root/reference/src/sqfa/_optim.py:90-96 -> model.py:190-220 / 508-546 ->
//  distances.py -> linalg.py and the autograd backward of all of it).
// It only sequences the kernels of project.cu and pairs.cu on the caller's stream with a
// caller-provided workspace, so one evaluation costs one host call instead of ~12.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/sqfa_b200.h"
#include "sqfa_internal.h"

namespace sqfa {

namespace {

__global__ void scale_loss_kernel(float* out, float w) { out[0] *= w; }

inline size_t al(size_t n) { return (n + 63) & ~size_t(63); }  // floats, 256-byte granules

struct ClosureLayout {
  size_t T, Psi, Mu, E, W, gE, gLog, gPsi, gMu, flag, proj, total;
};

ClosureLayout closure_layout(int C, int D, int k, int dist) {
  const int base = dist & 15;
  const int m = (base == SQFA_DIST_FISHER_RAO_LB) ? k + 1 : k;
  ClosureLayout L;
  size_t o = 0;
  L.T = o;    o += al((size_t)C * k * D);
  L.Psi = o;  o += al((size_t)C * k * k);
  L.Mu = o;   o += al((size_t)C * k);
  L.E = o;    o += al((size_t)C * m * m);
  L.W = o;    o += al((size_t)C * class_factor_floats(m, dist));
  L.gE = o;   o += al((size_t)C * m * m);
  L.gLog = o; o += al((size_t)C * m * m);
  L.gPsi = o; o += al((size_t)C * k * k);
  L.gMu = o;  o += al((size_t)C * k);
  L.flag = o; o += al(16);
  L.proj = o; o += al(project_workspace_bytes(C, D, k) / sizeof(float) + 1);
  L.total = o;
  return L;
}

}  // namespace

size_t fused_loss_workspace_bytes(int C, int D, int k, int dist) {
  return closure_layout(C, D, k, dist).total * sizeof(float);
}

cudaError_t launch_fused_loss(const float* S, const float* M, const float* F, int C, int D, int k, float noise,
                              int dist, int64_t pair_begin, int64_t pair_end, float* out, float* dF, float* ws,
                              cudaStream_t st) {
  const int base = dist & 15;
  const bool fr = base == SQFA_DIST_FISHER_RAO_LB;
  const bool le = base == SQFA_DIST_LOG_EUCLIDEAN;
  const int m = fr ? k + 1 : k;
  const ClosureLayout L = closure_layout(C, D, k, dist);
  float* T = ws + L.T;
  float* Psi = ws + L.Psi;
  float* Mu = ws + L.Mu;
  float* E = ws + L.E;
  float* W = ws + L.W;
  float* gE = ws + L.gE;
  float* gLog = ws + L.gLog;
  float* gPsi = ws + L.gPsi;
  float* gMu = ws + L.gMu;
  int32_t* flag = reinterpret_cast<int32_t*>(ws + L.flag);
  float* proj = ws + L.proj;
  const float* Mfr = fr ? M : nullptr;
  const int64_t P = (int64_t)C * (C - 1) / 2;
  const float weight = -1.0f / (float)(P > 0 ? P : 1);

  cudaError_t e;
  if ((e = launch_project_fwd(S, Mfr, F, C, D, k, T, Psi, fr ? Mu : nullptr, proj, st)) != cudaSuccess) return e;
  if ((e = launch_embed_fwd(Psi, Mu, noise, C, k, fr ? 1 : 0, E, st)) != cudaSuccess) return e;
  if ((e = launch_class_factor(E, C, m, dist, W, flag, st)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(out, 0, 2 * sizeof(float), st)) != cudaSuccess) return e;
  // gE and gLog are adjacent in the workspace: one memset clears both
  if ((e = cudaMemsetAsync(gE, 0, (L.gPsi - L.gE) * sizeof(float), st)) != cudaSuccess) return e;
  float* acc = le ? gLog : gE;
  if ((e = launch_pair_distances(W, W, C, C, m, dist, 1, pair_begin, pair_end, weight, nullptr, nullptr, out, acc,
                                 acc, nullptr, st)) != cudaSuccess)
    return e;
  if (le && (e = launch_class_factor_bwd(W, gLog, C, m, dist, gE, st)) != cudaSuccess) return e;
  if ((e = launch_embed_bwd(gE, Mu, C, k, fr ? 1 : 0, gPsi, gMu, st)) != cudaSuccess) return e;
  if ((e = launch_project_bwd(gPsi, fr ? gMu : nullptr, T, Mfr, C, D, k, dF, proj, st)) != cudaSuccess) return e;
  scale_loss_kernel<<<1, 1, 0, st>>>(out, weight);
  return cudaGetLastError();
}

}  // namespace sqfa
