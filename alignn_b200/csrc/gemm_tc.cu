// C[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ R[M,N]) in fp32 parity on the 5th-gen tensor cores:
// tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM) with the bf16x3 operand split (tc_common.cuh).
//
// This is the Linear-layer workhorse of the conv path: the four node projections (one
// [Nn,d]x[d,4d] GEMM), the edge gate (alignn.py:101) and the two data-gradient GEMMs of the
// backward.  A is the fp32 activation matrix, streamed from HBM once; it is converted to bf16 hi/lo
// planes by the loader warps on its way into shared memory (software-pipelined: the global loads of
// chunk k+1 are in flight while chunk k is converted).  W is pre-split once per step into an image
// that already has the UMMA core-matrix order (gemm_prepare_weights), so a K-chunk of it is ONE
// contiguous TMA bulk copy (cp.async.bulk -> UBLKCP) signalled on the stage's mbarrier.
//
// Persistent, warp-specialised CTA (one per SM), 128 x BN output tiles (BN = 256 when N allows):
//   warp 0      : TMEM owner + MMA issuer (one lane)
//   warps 1-4   : epilogue -- tcgen05.ld -> registers -> warp-private smem transpose -> coalesced
//                 128-byte row segments to HBM (+ bias, + residual)
//   warps 5-12  : loaders / fp32->bf16x2 converters
// 4-stage smem ring of BK=32 chunks (mbarrier full/empty), double-buffered TMEM accumulator
// (mbarrier tfull/tempty) so the epilogue of tile i overlaps the main loop of tile i+1.
#include "tc_common.cuh"
#include "api_common.h"
#include "alignn_b200.h"

namespace alignn {
namespace gemm {

// Development aid (tools/time_kernels.py): knock out one pipeline agent to find the limiter.
// bit0: no A loads, bit1: no W bulk copy, bit2: no C stores, bit3: no MMA.  0 in production.
static int g_debug_flags = 0;

constexpr int BM = 128;       // rows per tile (UMMA M)
constexpr int BK = 32;        // K per pipeline stage (2 UMMA K=16 steps)
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 4;
constexpr int LOAD_WARPS = 8;
constexpr int THREADS = 32 * (1 + EPI_WARPS + LOAD_WARPS);   // 416
constexpr uint32_t LBO = 128;               // next 8-element K chunk
constexpr uint32_t SBO = (BK / 8) * 128;    // next 8-row group (chunk-local image): 512 B
constexpr int EPI_COLS = 128;               // columns staged per epilogue pass (BN < 128: BN)
constexpr int EPI_STRIDE = EPI_COLS + 4;    // floats; padded staging row

template <int BN>
struct Cfg {
  static constexpr int A_PLANE = BM * BK * 2;   // bytes of one bf16 plane of the A chunk
  static constexpr int B_PLANE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_PLANE + 2 * B_PLANE;
  static constexpr int PIPE_BYTES = STAGES * STAGE;
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_STRIDE * 4;
  static constexpr int BAR_OFF = PIPE_BYTES + EPI_BYTES;
  static constexpr int SMEM = BAR_OFF + 128;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // double-buffered accumulator: 64..512
};

// byte offset of element (r, k) inside one chunk plane (rows x BK, core-matrix order)
__host__ __device__ constexpr int plane_off(int r, int k) { return (r >> 3) * (int)SBO + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }

// Loader thread -> (row, float4 index along K) of the 128 x 32 fp32 chunk for its i-th load.
// A half-warp (what one 64-bit shared store wavefront serves) covers 8 rows x 2 adjacent float4, i.e. one
// K-core-matrix column of 8 rows = 128 contiguous bytes of the plane: bank-conflict-free stores, and
// every lane pair still reads a full 32-byte sector from HBM.
__device__ __forceinline__ void a_coord(int i, int lt, int& row, int& kq) {
  const int w = lt >> 5, lane = lt & 31;
  const int u = i * LOAD_WARPS + w;                      // 32 units of (8 rows x 4 float4)
  row = (u >> 1) * 8 + ((lane >> 1) & 7);
  kq = (u & 1) * 4 + (lane >> 4) * 2 + (lane & 1);
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_nt_bf16x3_kernel(const float* __restrict__ A, int64_t lda, const uint8_t* __restrict__ Wimg, int M, int N, int K,
                      const float* __restrict__ bias, const float* __restrict__ R, int64_t ldr,
                      float* __restrict__ C, int64_t ldc, int dbg) {
  using F = Cfg<BN>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + F::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nk = K / BK;
  const int n_tiles = N / BN;
  const int m_tiles = (M + BM - 1) / BM;
  const int total = m_tiles * n_tiles;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full[s], LOAD_WARPS + 1); tc::mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tfull[a], 1); tc::mbar_init(&tempty[a], EPI_WARPS); }
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, F::TMEM_COLS);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 1 + EPI_WARPS) {
    // ================= loaders / converters =================
    const int lt = tid - 32 * (1 + EPI_WARPS);          // 0..255
    // The CTA's work is the stream of chunks c = (local tile, kc).  PF chunks of A are kept in flight in
    // registers.  The inner loop is kept lean on purpose: with only two loader warps per scheduler the
    // instruction stream is latency-bound, so every per-chunk address is an incremented pointer, not
    // recomputed index math (an earlier version spent ~3000 cycles per chunk on it).
    constexpr int PF = 3;
    float4 buf[PF][4];
    int soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row, kq;
      a_coord(i, lt, row, kq);
      soff[i] = plane_off(row, kq * 4);
    }
    const int my_tiles = (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nchunks = my_tiles * nk;
    // load cursor (runs PF chunks ahead of the store cursor)
    int l_tile = blockIdx.x, l_kc = 0;
    const float* lp[4];
    bool lval[4];
    auto set_tile_ptrs = [&](int tile) {
      const int m0 = (tile / n_tiles) * BM;
      // The epilogue adds the residual R row by row with only 4 warps: straight from HBM that costs ~1.5 us per
      // dependent load (the gy GEMM ran at 574 us instead of 150).  Pull this tile's R rows into L2 now, about
      // one tile ahead of the epilogue that will read them.
      if (R && lt < BM && m0 + lt < M)
        tc::bulk_prefetch_l2(R + (int64_t)(m0 + lt) * ldr + (tile % n_tiles) * BN, (uint32_t)BN * 4u);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r_, kq;
        a_coord(i, lt, r_, kq);
        const int gr = m0 + r_;
        lval[i] = (gr < M) && !(dbg & 1);
        lp[i] = A + (int64_t)(lval[i] ? gr : 0) * lda + kq * 4;
      }
    };
    auto load_next = [&](float4 (&v)[4]) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = lval[i] ? __ldcs(reinterpret_cast<const float4*>(lp[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
        lp[i] += BK;
      }
      if (++l_kc == nk) { l_kc = 0; l_tile += gridDim.x; if (l_tile < total) set_tile_ptrs(l_tile); }
    };
    if (nchunks > 0) set_tile_ptrs(l_tile);
#pragma unroll
    for (int j = 0; j < PF; ++j)
      if (j < nchunks) load_next(buf[j]);
    // store cursor
    int s = 0, ph = 0, s_kc = 0, s_ntile = (int)blockIdx.x % n_tiles, s_tile = blockIdx.x;
    const uint8_t* wsrc = Wimg + (int64_t)s_ntile * nk * 2 * F::B_PLANE;
    for (int c0 = 0; c0 < nchunks; c0 += PF) {
#pragma unroll
      for (int j = 0; j < PF; ++j) {
        const int c = c0 + j;
        if (c < nchunks) {
          if (c >= STAGES) { if (dbg & 512) tc::mbar_wait_poll(&empty[s], ph ^ 1); else tc::mbar_wait(&empty[s], ph ^ 1); }
          uint8_t* st = smem + s * F::STAGE;
          if (lt == 0) {   // W chunk: one contiguous bulk copy (both planes), counted in bytes on full[s]
            if (dbg & 2) {
              tc::mbar_arrive(&full[s]);
            } else {
              tc::mbar_arrive_expect_tx(&full[s], 2 * F::B_PLANE);
              tc::bulk_g2s(st + 2 * F::A_PLANE, wsrc, 2 * F::B_PLANE, &full[s]);
            }
          }
          if (!(dbg & 16)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint2 hi, lo;
              tc::split4(buf[j][i], hi, lo);
              *reinterpret_cast<uint2*>(st + soff[i]) = hi;
              *reinterpret_cast<uint2*>(st + F::A_PLANE + soff[i]) = lo;
            }
          }
          if (c + PF < nchunks) load_next(buf[j]);           // refill this register slot
          tc::fence_async_smem();
          __syncwarp();
          if ((lt & 31) == 0) { if (dbg & 1024) tc::mbar_arrive_relaxed(&full[s]); else tc::mbar_arrive(&full[s]); }  // one arrival per warp
          // advance the store cursor
          wsrc += 2 * F::B_PLANE;
          if (++s_kc == nk) {
            s_kc = 0; s_tile += gridDim.x; s_ntile = s_tile % n_tiles;
            wsrc = Wimg + (int64_t)s_ntile * nk * 2 * F::B_PLANE;
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 1) {
    // ================= epilogue =================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem + F::PIPE_BYTES) + (warp - 1) * 32 * EPI_STRIDE;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
      tc::mbar_wait(&tfull[acc], (lt >> 1) & 1);
      tc::fence_after_sync();
      // EC columns at a time: TMEM -> registers -> warp-private staging (thread per row) -> the warp writes one
      // row's EC contiguous floats per instruction (512 B for EC = 128): long contiguous runs for DRAM.
      constexpr int EC = BN < EPI_COLS ? BN : EPI_COLS;
#pragma unroll 1
      for (int c0 = 0; c0 < ((dbg & 32) ? 0 : BN); c0 += EC) {
#pragma unroll 1
        for (int cc = 0; cc < EC; cc += 32) {
          float v[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0 + cc), v);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * EPI_STRIDE + cc + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        constexpr int LPR = EC / 4;                          // lanes per row (float4 each): 32, 16 or 8
        constexpr int RPI = 32 / LPR;                        // rows per store instruction
        const int c4 = (lane % LPR) * 4;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c0 + c4));
        constexpr int RB = 8;                                // rows per batch: all residual loads of a batch first
#pragma unroll 1
        for (int rb = 0; rb < 32; rb += RB * RPI) {
          float4 qv[RB];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int gr = m0 + q * 32 + rb + u * RPI + lane / LPR;
            qv[u] = (R && gr < M) ? __ldcs(reinterpret_cast<const float4*>(R + (int64_t)gr * ldr + n0 + c0 + c4))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int r = rb + u * RPI + lane / LPR;
            const int gr = m0 + q * 32 + r;
            float4 o = *reinterpret_cast<const float4*>(stg + r * EPI_STRIDE + c4);
            o.x += b4.x + qv[u].x; o.y += b4.y + qv[u].y; o.z += b4.z + qv[u].z; o.w += b4.w + qv[u].w;
            if (gr < M && !(dbg & 4)) *reinterpret_cast<float4*>(C + (int64_t)gr * ldc + n0 + c0 + c4) = o;
          }
        }
        __syncwarp();
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tempty[acc]);      // this accumulator buffer may be overwritten
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(BM, BN);
    // descriptors differ only in their start-address field (bits [0,14), units of 16 bytes)
    const uint64_t desc0 = tc::smem_desc(tc::smem_u32(smem), LBO, SBO);
    uint32_t lt = 0;
    int s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      if (lt >= 2) tc::mbar_wait(&tempty[acc], ((lt >> 1) - 1) & 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem + (uint32_t)(acc * BN);
      uint32_t accum = 0;
      for (int kc = 0; kc < nk; ++kc) {
        if (dbg & 512) tc::mbar_wait_poll(&full[s], ph); else tc::mbar_wait(&full[s], ph);
        tc::fence_after_sync();
        const uint64_t sd = desc0 + (uint64_t)((s * F::STAGE) >> 4);
        if (!(dbg & 8)) {
#pragma unroll
          for (int j = 0; j < BK / 16; ++j) {
            const uint64_t a_hi = sd + (uint64_t)((j * 2 * LBO) >> 4);
            const uint64_t a_lo = a_hi + (uint64_t)(F::A_PLANE >> 4);
            const uint64_t b_hi = a_hi + (uint64_t)((2 * F::A_PLANE) >> 4);
            const uint64_t b_lo = b_hi + (uint64_t)(F::B_PLANE >> 4);
            tc::mma_bf16_ss(d_tmem, a_lo, b_hi, IDESC, accum);   // small terms first
            tc::mma_bf16_ss(d_tmem, a_hi, b_lo, IDESC, 1);
            tc::mma_bf16_ss(d_tmem, a_hi, b_hi, IDESC, 1);
            accum = 1;
          }
        }
        if (dbg & 256) tc::mbar_arrive(&empty[s]); else
        tc::mma_commit(&empty[s]);                       // frees the stage when these MMAs retire
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      tc::mma_commit(&tfull[acc]);                       // accumulator complete -> epilogue
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, F::TMEM_COLS);
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

// Table-driven variant: every entry converts one source block into its place inside a (possibly larger) image, so
// all operand images of a model -- [W_sg; W_du; W_dg; W_su] stacked along N, its transpose stacked along K, the
// edge gate and its transpose, the embedding Linears (K zero-padded) -- are refreshed by ONE launch per step.
// blockIdx.y = entry; thread = one 8-element K group of one image row inside the entry's block.
__global__ void prepare_weights_table_kernel(const alignn_b200_image_entry* __restrict__ entries) {
  const alignn_b200_image_entry e = entries[blockIdx.y];
  const int Ne = e.transpose ? e.cols : e.rows, Ke = e.transpose ? e.rows : e.cols;   // block shape in image coordinates
  const int k8n = (Ke + 7) / 8;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)Ne * k8n) return;
  const int nl = (int)(t / k8n), k8 = (int)(t % k8n);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int kl = k8 * 8 + j;
    v[j] = kl < Ke ? (e.transpose ? e.W[(int64_t)kl * e.ldw + nl] : e.W[(int64_t)nl * e.ldw + kl]) : 0.f;
  }
  uint2 h0, l0, h1, l1;
  tc::split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
  tc::split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
  const int bn = (e.N % 256 == 0) ? 256 : (e.N % 128 == 0) ? 128 : (e.N % 64 == 0) ? 64 : 32;
  const int n = e.n_off + nl, k = e.k_off + k8 * 8;          // image coordinates (k_off is a multiple of 8)
  const int nt = n / bn, r = n % bn, kc = k / BK, kk = (k % BK) / 8;
  const int64_t b_plane = (int64_t)bn * BK * 2;
  const int64_t chunk = ((int64_t)nt * (e.K / BK) + kc) * 2 * b_plane;
  const int off = plane_off(r, kk * 8);
  uint8_t* img = reinterpret_cast<uint8_t*>(e.image);
  *reinterpret_cast<uint4*>(img + chunk + off) = make_uint4(h0.x, h0.y, h1.x, h1.y);
  *reinterpret_cast<uint4*>(img + chunk + b_plane + off) = make_uint4(l0.x, l0.y, l1.x, l1.y);
}

// dst[j] = a[j] (+ b[j]): the stacked / folded bias vectors that go with the images (blockIdx.y = entry)
__global__ void prepare_bias_table_kernel(const alignn_b200_bias_entry* __restrict__ entries) {
  const alignn_b200_bias_entry e = entries[blockIdx.y];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < e.n; j += gridDim.x * blockDim.x)
    e.dst[j] = e.a[j] + (e.b ? e.b[j] : 0.f);
}

inline int pick_bn(int N) { return (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : (N % 32 == 0) ? 32 : 0; }

template <int BN>
int launch_gemm(const float* A, int64_t lda, const void* img, int M, int N, int K, const float* bias, const float* R,
                int64_t ldr, float* C, int64_t ldc, cudaStream_t st) {
  using F = Cfg<BN>;
  static alignn::DeviceOnce configured; int cfg_dev;   // idempotent attribute; a benign race sets it twice
  if (configured.needed(&cfg_dev)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_bf16x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM);
    if (e != cudaSuccess) return record_cuda_error((int)e);
    configured.done(cfg_dev);
  }
  const int total = ((M + BM - 1) / BM) * (N / BN);
  const int grid = total < 148 ? total : 148;          // persistent: one CTA per SM
  gemm_nt_bf16x3_kernel<BN><<<grid, THREADS, F::SMEM, st>>>(A, lda, reinterpret_cast<const uint8_t*>(img), M, N, K, bias, R,
                                                           ldr, C, ldc, g_debug_flags);
  return check_launch();
}

}  // namespace gemm
}  // namespace alignn

extern "C" {

void alignn_b200_debug_gemm_flags(int flags) { alignn::gemm::g_debug_flags = flags; }   /* not in the public header */

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
  if (bn == 256) prepare_weights_kernel<256><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  else if (bn == 128) prepare_weights_kernel<128><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  else if (bn == 64) prepare_weights_kernel<64><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  else prepare_weights_kernel<32><<<blocks, 256, 0, st>>>(W, N, K, ldw, transpose, img);
  return alignn::check_launch();
}

int alignn_b200_gemm_prepare_table(const alignn_b200_image_entry* entries, int n_entries, int64_t max_units,
                                   const alignn_b200_bias_entry* bias_entries, int n_bias, alignn_stream_t stream) {
  if (n_entries < 0 || n_bias < 0 || (n_entries > 0 && (!entries || max_units <= 0)) || (n_bias > 0 && !bias_entries))
    return ALIGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_entries > 0) {
    const unsigned bx = (unsigned)((max_units + 255) / 256);
    alignn::gemm::prepare_weights_table_kernel<<<dim3(bx, (unsigned)n_entries), 256, 0, st>>>(entries);
    int rc = alignn::check_launch();
    if (rc != ALIGNN_OK) return rc;
  }
  if (n_bias > 0) {
    alignn::gemm::prepare_bias_table_kernel<<<dim3(4, (unsigned)n_bias), 256, 0, st>>>(bias_entries);
    return alignn::check_launch();
  }
  return ALIGNN_OK;
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
    case 256: return launch_gemm<256>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    case 128: return launch_gemm<128>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    case 64: return launch_gemm<64>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    case 32: return launch_gemm<32>(A, lda, w_image, (int)M, N, K, bias, R, ldr, C, ldc, st);
    default: return ALIGNN_ERR_UNSUPPORTED_D;
  }
}

}  // extern "C"
