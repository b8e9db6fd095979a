// STAGED WORK (see egc_fused.h): the node side that follows the fused edge kernel when the norm is a LayerNorm:
//   out = (res ? res : 0) + silu(LayerNorm(R) * gamma + beta)        rows [n, d]   (alignn_atomwise.py:209-211)
// One warp per row, two rows in flight, the same row helpers (and therefore the same arithmetic) as the node
// finalisation inside the shipped egc_forward_kernel.  The BatchNorm variants need nothing new:
// alignn_b200_rowstats_partials + alignn_b200_bn_finalize + alignn_b200_affine_silu_residual already exist.
#include "../common.cuh"
#include "alignn_b200.h"
#include "egc_fused.h"

namespace alignn {
namespace staged {

__device__ __forceinline__ float silu_(float u) { return u * sigmoidf_(u); }

template <int D>
__global__ void __launch_bounds__(kThreads)
ln_silu_residual_kernel(const float* __restrict__ R, const float* __restrict__ res, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, float* __restrict__ out, int64_t n) {
  constexpr int V = RowCfg<D>::VPL;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float w[V], b[V];
  ld_vec<D>(w, gamma, lane);
  ld_vec<D>(b, beta, lane);
  for (int64_t r = warp0; r < n; r += 2 * nwarps) {
    const int64_t r2 = r + nwarps;
    const bool has2 = r2 < n;
    float v[V], v2[V], y[V], y2[V];
    ld_row<D, true>(v, R + r * D, lane);
    if (has2) ld_row<D, true>(v2, R + r2 * D, lane);
    if (res) {
      ld_row<D, true>(y, res + r * D, lane);
      if (has2) ld_row<D, true>(y2, res + r2 * D, lane);
    }
    float mean, rstd;
    row_mean_rstd<D>(v, eps, mean, rstd);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = silu_((v[k] - mean) * rstd * w[k] + b[k]) + (res ? y[k] : 0.f);
    st_row<D, true>(out + r * D, v, lane);
    if (has2) {
      row_mean_rstd<D>(v2, eps, mean, rstd);
#pragma unroll
      for (int k = 0; k < V; ++k) v2[k] = silu_((v2[k] - mean) * rstd * w[k] + b[k]) + (res ? y2[k] : 0.f);
      st_row<D, true>(out + r2 * D, v2, lane);
    }
  }
}

// Node side of the train-mode BatchNorm backward (the node block of egc_backward_dst_kernel on its own): one warp
// per node row.  gx' = scale * (gu - c1 - xhat * c2) with gu = gx_out * silu'(x' * scale + shift); the two factors the
// edge side multiplies with: GSh = gx' / (S + eps), GS = -gx' * h / (S + eps).
__device__ __forceinline__ float dsilu_(float u) { const float s = sigmoidf_(u); return s * (1.f + u * (1.f - s)); }

template <int D>
__global__ void __launch_bounds__(kThreads)
egc_backward_nodes_kernel(const float* __restrict__ XP, const float* __restrict__ gx_out, const float* __restrict__ S,
                          const float* __restrict__ H, const float* __restrict__ n_w, const float* __restrict__ n_b,
                          const float* __restrict__ n_mean, const float* __restrict__ n_rstd, const float* __restrict__ n_c1,
                          const float* __restrict__ n_c2, float gate_eps, int64_t Nn, float* __restrict__ GPD, int64_t ld,
                          float* __restrict__ GSh, float* __restrict__ GS, float* __restrict__ partials) {
  constexpr int V = RowCfg<D>::VPL;
  __shared__ float red[kWarpsPerBlock * D];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float w[V], b[V], mu[V], rs[V], c1[V], c2[V], acc[1][V];
  ld_vec<D>(w, n_w, lane); ld_vec<D>(b, n_b, lane); ld_vec<D>(mu, n_mean, lane);
  ld_vec<D>(rs, n_rstd, lane); ld_vec<D>(c1, n_c1, lane); ld_vec<D>(c2, n_c2, lane);
#pragma unroll
  for (int k = 0; k < V; ++k) acc[0][k] = 0.f;
  for (int64_t v = warp0; v < Nn; v += nwarps) {
    float xp[V], go[V], sv[V], hv[V], gsh[V], gs[V];
    ld_row<D, false>(xp, XP + v * D, lane);
    ld_row<D, false>(go, gx_out + v * D, lane);
    ld_row<D, false>(sv, S + v * D, lane);
    ld_row<D, false>(hv, H + v * D, lane);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float gu = go[k] * dsilu_(xp[k] * w[k] + b[k]);
      const float xh = (xp[k] - mu[k]) * rs[k];
      float gr = w[k] * gu;
      gr -= w[k] * (c1[k] + xh * c2[k]);
      const float inv = 1.f / (sv[k] + gate_eps);
      gsh[k] = gr * inv;
      gs[k] = -gr * hv[k] * inv;
      xp[k] = gr;
      acc[0][k] += gr;
    }
    st_row<D, false>(GPD + v * ld, xp, lane);
    st_row<D, false>(GSh + v * D, gsh, lane);
    st_row<D, false>(GS + v * D, gs, lane);
  }
  if (partials) block_reduce_to_partials<D, 1>(acc, partials + (int64_t)blockIdx.x * D, red);
}

}  // namespace staged
}  // namespace alignn

extern "C" int alignn_b200_egc_backward_nodes(const float* XP, const float* gx_out, const float* S, const float* H,
                                              const float* n_w, const float* n_b, const float* n_mean, const float* n_rstd,
                                              const float* n_c1, const float* n_c2, float gate_eps, int64_t Nn, int d,
                                              float* GPD, int64_t ld_gpd, float* GSh, float* GS, float* partials,
                                              int partial_rows, void* stream) {
  using namespace alignn;
  if (d != 32 && d != 64 && d != 128 && d != 256) return ALIGNN_ERR_UNSUPPORTED_D;
  if (Nn < 0 || ld_gpd < d || (ld_gpd % 4)) return ALIGNN_ERR_BAD_ARG;
  if (Nn == 0) return ALIGNN_OK;
  if (!XP || !gx_out || !S || !H || !n_w || !n_b || !n_mean || !n_rstd || !n_c1 || !n_c2 || !GPD || !GSh || !GS)
    return ALIGNN_ERR_BAD_ARG;
  int64_t blocks = (Nn + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  if (partials && partial_rows != (int)blocks) return ALIGNN_ERR_WORKSPACE;   // one partial row per block
  cudaStream_t st = (cudaStream_t)stream;
#define ALIGNN_NODES(DD) staged::egc_backward_nodes_kernel<DD><<<(int)blocks, kThreads, 0, st>>>( \
      XP, gx_out, S, H, n_w, n_b, n_mean, n_rstd, n_c1, n_c2, gate_eps, Nn, GPD, ld_gpd, GSh, GS, partials)
  switch (d) {
    case 256: ALIGNN_NODES(256); break;
    case 128: ALIGNN_NODES(128); break;
    case 64: ALIGNN_NODES(64); break;
    default: ALIGNN_NODES(32); break;
  }
#undef ALIGNN_NODES
  return cudaGetLastError() == cudaSuccess ? ALIGNN_OK : ALIGNN_ERR_CUDA;
}

extern "C" int alignn_b200_ln_silu_residual(const float* R, const float* res, const float* gamma, const float* beta, float eps,
                                            float* out, int64_t n, int d, void* stream) {
  using namespace alignn;
  if (d != 32 && d != 64 && d != 128 && d != 256) return ALIGNN_ERR_UNSUPPORTED_D;
  if (n < 0 || (n > 0 && (!R || !gamma || !beta || !out))) return ALIGNN_ERR_BAD_ARG;
  if (n == 0) return ALIGNN_OK;
  int64_t blocks = (n + 2 * kWarpsPerBlock - 1) / (2 * kWarpsPerBlock);
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 256: staged::ln_silu_residual_kernel<256><<<(int)blocks, kThreads, 0, st>>>(R, res, gamma, beta, eps, out, n); break;
    case 128: staged::ln_silu_residual_kernel<128><<<(int)blocks, kThreads, 0, st>>>(R, res, gamma, beta, eps, out, n); break;
    case 64: staged::ln_silu_residual_kernel<64><<<(int)blocks, kThreads, 0, st>>>(R, res, gamma, beta, eps, out, n); break;
    default: staged::ln_silu_residual_kernel<32><<<(int)blocks, kThreads, 0, st>>>(R, res, gamma, beta, eps, out, n); break;
  }
  return cudaGetLastError() == cudaSuccess ? ALIGNN_OK : ALIGNN_ERR_CUDA;
}
