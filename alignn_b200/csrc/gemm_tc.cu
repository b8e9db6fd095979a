// C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ R[M,N]) in fp32 parity on the 5th-gen tensor cores:
// tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM) with the bf16x3 operand split (tc_common.cuh).
//
// This is the Linear-layer workhorse of the conv path: the four node projections (one
// [Nn,d]x[d,4d] GEMM), the edge gate (alignn.py:101) and the two data-gradient GEMMs of the
// backward.  A is the fp32 activation matrix streamed from HBM exactly once per N tile; it is
// converted to bf16 hi/lo planes by the loader threads on its way into shared memory.  W is
// pre-split once per step into an image that already has the UMMA core-matrix order
// (gemm_prepare_weights), so a tile of it is a plain 16-byte-vector copy.
//
// One CTA = one 128 x BN output tile; 2 CTAs per SM so that one CTA's epilogue overlaps the other's
// main loop.  Warps 0-3: loader/converter, then epilogue (TMEM -> registers -> smem transpose ->
// coalesced row stores).  Warp 4, one lane: MMA issuer.  3-stage smem ring of BK=32 chunks,
// mbarrier full/empty pairs; the accumulator lives in TMEM (BN columns).
#include "tc_common.cuh"
#include "api_common.h"
#include "alignn_b200.h"

namespace alignn {
namespace gemm {

constexpr int BM = 128;       // rows per CTA tile  (UMMA M)
constexpr int BK = 32;        // K per pipeline stage (2 UMMA K=16 steps)
constexpr int STAGES = 3;
constexpr int LOADERS = 128;  // threads in warps 0-3
constexpr int THREADS = 160;  // + MMA warp
constexpr uint32_t LBO = 128;               // next 8-element K chunk
constexpr uint32_t SBO = (BK / 8) * 128;    // next 8-row group (chunk-local image): 512 B

template <int BN>
struct Cfg {
  static constexpr int A_PLANE = BM * BK * 2;   // bytes of one bf16 plane of the A chunk
  static constexpr int B_PLANE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int CS_STRIDE = BN + 4;      // floats; padded epilogue staging row
  static constexpr int CS_BYTES = BM * CS_STRIDE * 4;
  static constexpr int PIPE_BYTES = STAGES * STAGE;
  static_assert(CS_BYTES <= PIPE_BYTES, "epilogue staging must fit in the drained pipeline buffers");
  static constexpr int BAR_OFF = PIPE_BYTES;
  static constexpr int SMEM = PIPE_BYTES + 128;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// byte offset of element (r, k) inside one chunk plane (rows x BK, core-matrix order)
__host__ __device__ constexpr int plane_off(int r, int k) { return (r >> 3) * (int)SBO + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

template <int BN>
__global__ void __launch_bounds__(THREADS, 2)
gemm_nt_bf16x3_kernel(const float* __restrict__ A, int64_t lda, const uint4* __restrict__ Wimg, int M, int K,
                      const float* __restrict__ bias, const float* __restrict__ R, int64_t ldr,
                      float* __restrict__ C, int64_t ldc) {
  using F = Cfg<BN>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + F::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* accbar = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accbar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n_tile = blockIdx.y, n0 = n_tile * BN;
  const int nk = K / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], LOADERS); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(accbar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 4) tc::tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ================= loader / converter =================
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % STAGES;
      if (kc >= STAGES) tc::mbar_wait(&empty[s], ((kc / STAGES) - 1) & 1);
      uint8_t* st = smem + s * F::STAGE;
      // A chunk: 128 rows x 32 fp32 = 1024 float4; 8 lanes cover one row's 128 contiguous bytes
      float4 v[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int f = it * LOADERS + tid, row = f >> 3, kq = f & 7;
        const int gr = m0 + row;
        v[it] = (gr < M) ? __ldcs(reinterpret_cast<const float4*>(A + (int64_t)gr * lda + kc * BK + kq * 4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      // W chunk image: 2 planes, already in core-matrix order
      const uint4* wsrc = Wimg + ((int64_t)(n_tile * nk + kc) * 2 * F::B_PLANE) / 16;
      uint4* wdst = reinterpret_cast<uint4*>(st + 2 * F::A_PLANE);
      constexpr int WV = 2 * F::B_PLANE / 16;
      uint4 wv[(WV + LOADERS - 1) / LOADERS];
#pragma unroll
      for (int i = 0; i < (WV + LOADERS - 1) / LOADERS; ++i) {
        const int idx = i * LOADERS + tid;
        if (idx < WV) wv[i] = __ldg(wsrc + idx);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int f = it * LOADERS + tid, row = f >> 3, kq = f & 7;
        uint2 hi, lo;
        tc::split4(v[it], hi, lo);
        const int off = plane_off(row, kq * 4);
        *reinterpret_cast<uint2*>(st + off) = hi;
        *reinterpret_cast<uint2*>(st + F::A_PLANE + off) = lo;
      }
#pragma unroll
      for (int i = 0; i < (WV + LOADERS - 1) / LOADERS; ++i) {
        const int idx = i * LOADERS + tid;
        if (idx < WV) wdst[idx] = wv[i];
      }
      tc::fence_async_smem();
      tc::mbar_arrive(&full[s]);
    }
    // ================= epilogue =================
    tc::mbar_wait(accbar, 0);
    tc::fence_after_sync();
    float* Cs = reinterpret_cast<float*>(smem);          // pipeline buffers are drained by now
    const int row = warp * 32 + lane;                    // TMEM lane == tile row
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(Cs + row * F::CS_STRIDE + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the 4 epilogue warps only
    // coalesced row stores (+ bias, + residual)
    constexpr int V4_PER_ROW = BN / 4;
    for (int idx = tid; idx < BM * V4_PER_ROW; idx += LOADERS) {
      const int r = idx / V4_PER_ROW, c = (idx % V4_PER_ROW) * 4;
      const int gr = m0 + r;
      if (gr >= M) continue;
      float4 o = *reinterpret_cast<const float4*>(Cs + r * F::CS_STRIDE + c);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0 + c));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      if (R) {
        const float4 q = __ldcs(reinterpret_cast<const float4*>(R + (int64_t)gr * ldr + n0 + c));
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
      }
      *reinterpret_cast<float4*>(C + (int64_t)gr * ldc + n0 + c) = o;
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(BM, BN);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % STAGES;
      tc::mbar_wait(&full[s], (kc / STAGES) & 1);
      tc::fence_after_sync();
      const uint32_t base = tc::smem_u32(smem + s * F::STAGE);
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        const uint32_t ko = j * 2 * LBO;               // two core matrices per K=16 step
        const uint64_t a_hi = tc::smem_desc(base + ko, LBO, SBO);
        const uint64_t a_lo = tc::smem_desc(base + F::A_PLANE + ko, LBO, SBO);
        const uint64_t b_hi = tc::smem_desc(base + 2 * F::A_PLANE + ko, LBO, SBO);
        const uint64_t b_lo = tc::smem_desc(base + 2 * F::A_PLANE + F::B_PLANE + ko, LBO, SBO);
        tc::mma_bf16_ss(tmem, a_lo, b_hi, IDESC, (kc | j) != 0);   // small terms first
        tc::mma_bf16_ss(tmem, a_hi, b_lo, IDESC, 1);
        tc::mma_bf16_ss(tmem, a_hi, b_hi, IDESC, 1);
      }
      tc::mma_commit(&empty[s]);                       // frees the stage when these MMAs retire
    }
    tc::mma_commit(accbar);                            // accumulator complete
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, F::TMEM_COLS);
}

// W[N,K] fp32 (row stride ldw; or, if transpose, the N x K matrix is W^T of a [K,N] array) ->
// image [N/BN tiles][K/32 chunks][hi, lo][BN x 32 bf16 in core-matrix order]
template <int BN>
__global__ void prepare_weights_kernel(const float* __restrict__ W, int N, int K, int64_t ldw, int transpose,
                                       uint8_t* __restrict__ img) {
  using F = Cfg<BN>;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 8-element K group of one row
  const int k8n = K / 8;
  if (t >= (int64_t)N * k8n) return;
  const int n = (int)(t / k8n), k8 = (int)(t % k8n);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = k8 * 8 + j;
    v[j] = transpose ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
  }
  uint2 h0, l0, h1, l1;
  tc::split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
  tc::split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
  const int nt = n / BN, r = n % BN, kc = k8 / (BK / 8), kk = k8 % (BK / 8);
  const int64_t chunk = ((int64_t)nt * (K / BK) + kc) * 2 * F::B_PLANE;
  const int off = plane_off(r, kk * 8);
  *reinterpret_cast<uint4*>(img + chunk + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(img + chunk + F::B_PLANE + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

inline int pick_bn(int N) { return (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : (N % 32 == 0) ? 32 : 0; }

template <int BN>
int launch_gemm(const float* A, int64_t lda, const void* img, int M, int N, int K, const float* bias, const float* R,
                int64_t ldr, float* C, int64_t ldc, cudaStream_t st) {
  using F = Cfg<BN>;
  static bool configured = false;   // idempotent attribute; a benign race sets it twice
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_bf16x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM);
    if (e != cudaSuccess) return record_cuda_error((int)e);
    configured = true;
  }
  dim3 grid((M + BM - 1) / BM, N / BN);
  gemm_nt_bf16x3_kernel<BN><<<grid, THREADS, F::SMEM, st>>>(A, lda, reinterpret_cast<const uint4*>(img), M, K, bias, R, ldr,
                                                           C, ldc);
  return check_launch();
}

}  // namespace gemm
}  // namespace alignn

extern "C" {

size_t alignn_b200_gemm_weight_image_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || alignn::gemm::pick_bn(N) == 0 || K % alignn::gemm::BK != 0) return 0;
  return (size_t)N * K * 2 * 2;   // two bf16 planes
}

int alignn_b200_gemm_prepare_weights(const float* W, int N, int K, int64_t ldw, int transpose, void* image,
                                     alignn_stream_t stream) {
  using namespace alignn::gemm;
  if (!W || !image || N <= 0 || K <= 0 || K % BK != 0) return ALIGNN_ERR_BAD_ARG;
  const int bn = pick_bn(N);
  if (bn == 0) return ALIGNN_ERR_UNSUPPORTED_D;
  const int64_t total = (int64_t)N * (K / 8);
  const int blocks = (int)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* img = reinterpret_cast<uint8_t*>(image);
  if (bn == 128) prepare_weights_kernel<128><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  else if (bn == 64) prepare_weights_kernel<64><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  else prepare_weights_kernel<32><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  return alignn::check_launch();
}

int alignn_b200_gemm_nt(const float* A, int64_t lda, const void* w_image, int64_t M, int N, int K, const float* bias,
                        const float* R, int64_t ldr, float* C, int64_t ldc, alignn_stream_t stream) {
  using namespace alignn::gemm;
  if (M < 0 || N <= 0 || K <= 0 || K % BK != 0 || lda < K || ldc < N || (R && ldr < N)) return ALIGNN_ERR_BAD_ARG;
  if (M == 0) return ALIGNN_OK;
  if (!A || !w_image || !C || M > 0x7fffffff) return ALIGNN_ERR_BAD_ARG;
  if ((lda % 4) || (ldc % 4) || (R && (ldr % 4))) return ALIGNN_ERR_BAD_ARG;   // 16-byte row alignment
  const int bn = pick_bn(N);
  cudaStream_t st = (cudaStream_t)stream;
  switch (bn) {
    case 128: return launch_gemm<128>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    case 64: return launch_gemm<64>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    case 32: return launch_gemm<32>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    default: return ALIGNN_ERR_UNSUPPORTED_D;
  }
}

}  // extern "C"
