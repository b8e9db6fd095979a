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

}  // namespace staged
}  // namespace alignn

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
