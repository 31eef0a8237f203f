// Weight gradient of the typed linear on the 5th-gen tensor cores:
//
//   dW[t] = dY[rows of type t]^T . X[rows of type t]            [T, M, Nn]   (M = n_out, Nn = K of the forward)
//
// the `loss.backward()` product behind every per-node-type nn.Linear of the reference (trainer/train_gnn.py:68-71 through
// models/HEATNet4.py:100-102,134,202 / models/HGT.py:82-84,121).  The reduction runs over the ROWS of a node type, so both
// operands are read "transposed".  Nothing is transposed in memory: dY and X stay in the [hi; lo] bf16 operand form the
// forward / data-gradient GEMMs already consume ([2N, cols] row-major), and the tensor core reads them MN-major - a TMA
// box of 64 rows x 64 columns (128 B per row, 128 B swizzle) IS the canonical MN-major SWIZZLE_128B atom sequence
// (8 reduction rows x 64 contiguous M/N elements per 1024 B atom; atoms along the reduction 1024 B apart = SBO, the
// next 64 M/N elements one box = 8192 B apart = LBO), instruction descriptor a_major = b_major = 1.
//
//   typed_wgrad_tc_kernel   persistent CTA pairs (cta_group::2), 256 x 256 fp32 accumulator in TMEM, 3-stage ring of
//                           64 KB stages (A_hi, A_lo, B_hi, B_lo: 2 boxes each), 3 MMAs per 16-row slice
//                           (hi.hi + hi.lo + lo.hi).  The rows of a type are cut into CHUNKS of whole 64-row blocks so
//                           that chunks x output tiles fill the 74 pairs (~2 tiles each); every (chunk, tile) writes a
//                           partial product with TMA stores.
//   wgrad_reduce_kernel     dW[t] = sum of the type's chunk partials + the < 64 tail rows of the type (which no 64-row
//                           block may cover: the next rows belong to another type) as a small fp32 SIMT product.
// Tensor-pipe bound: 2 * N * M * Nn algorithmic flops (x3 issued); operands come from L2 / HBM once per output tile column.
#include <cuda.h>

#include <mutex>

#include "tc_ptx.cuh"

namespace {

constexpr int BK = 64;                       // reduction rows per k-block
constexpr int UMMA_K = 16;
constexpr int BOX_BYTES = 64 * 64 * 2;       // one TMA box: 64 rows x 64 bf16
constexpr int OPER_BYTES = 2 * BOX_BYTES;    // 128 M/N elements of one operand plane
constexpr int STAGE_BYTES = 4 * OPER_BYTES;  // A_hi, A_lo, B_hi, B_lo
constexpr int STAGES = 3;
constexpr int EPI_WARPS = 8;
constexpr int EPI_WARP_BYTES = 32 * 32 * 4;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_WARPS * EPI_WARP_BYTES + 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_CHUNKS = 192;

struct WgradChunks {
  int n;
  int row0[MAX_CHUNKS];      // first row of the chunk
  int nblk[MAX_CHUNKS];      // whole 64-row blocks in it
};

struct WgradArgs {
  int n_rows;                // N: row offset of the lo plane of both operands
  int M, Nn;
  int n_tm, n_tn;            // 256 x 256 output tiles
  uint32_t idesc;
};

// MN-major, 128B-swizzled operand: LBO = distance between 64-element M/N groups (one box), SBO = distance between
// 8-row groups of the reduction dimension (one swizzle atom)
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
typed_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmP, const __grid_constant__ WgradChunks ch, WgradArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t epi_base = base + STAGES * STAGE_BYTES;
  const uint32_t bars = epi_base + EPI_WARPS * EPI_WARP_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES,
                 tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint16_t pair_mask = 3;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_per_chunk = a.n_tm * a.n_tn;
  const int total_tiles = ch.n * tiles_per_chunk;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  wsi_pdl_trigger();
  wsi_pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer (one lane per CTA)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int b = tile / tiles_per_chunk, rem = tile - b * tiles_per_chunk;
        const int tm = rem / a.n_tn, tn = rem - tm * a.n_tn;
        const int m0 = tm * 256 + (int)rank * 128, n0 = tn * 256 + (int)rank * 128;
        const int r0 = ch.row0[b], nb = ch.nblk[b];
        for (int kb = 0; kb < nb; ++kb) {
          const int r = r0 + kb * BK;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = mapa(full_bar + 8 * stage, 0);
          const uint32_t s0 = base + stage * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar + 8 * stage, 2 * STAGE_BYTES);
          tma_load_2d(&tmA, fb, s0, m0, r);
          tma_load_2d(&tmA, fb, s0 + BOX_BYTES, m0 + 64, r);
          tma_load_2d(&tmA, fb, s0 + OPER_BYTES, m0, a.n_rows + r);
          tma_load_2d(&tmA, fb, s0 + OPER_BYTES + BOX_BYTES, m0 + 64, a.n_rows + r);
          tma_load_2d(&tmB, fb, s0 + 2 * OPER_BYTES, n0, r);
          tma_load_2d(&tmB, fb, s0 + 2 * OPER_BYTES + BOX_BYTES, n0 + 64, r);
          tma_load_2d(&tmB, fb, s0 + 3 * OPER_BYTES, n0, a.n_rows + r);
          tma_load_2d(&tmB, fb, s0 + 3 * OPER_BYTES + BOX_BYTES, n0 + 64, a.n_rows + r);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA, one elected lane)
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t lbo = BOX_BYTES, sbo = 1024;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int b = tile / tiles_per_chunk;
        const int nb = ch.nblk[b];
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256;
        for (int kb = 0; kb < nb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t s0 = base + stage * STAGE_BYTES;
            const uint64_t a_hi = make_mn_desc(s0, lbo, sbo), a_lo = make_mn_desc(s0 + OPER_BYTES, lbo, sbo);
            const uint64_t b_hi = make_mn_desc(s0 + 2 * OPER_BYTES, lbo, sbo), b_lo = make_mn_desc(s0 + 3 * OPER_BYTES, lbo, sbo);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 128) >> 4);   // 16 reduction rows = 2 swizzle atoms = 2048 B
              tc_mma_f16(d_tmem, a_hi + adv, b_hi + adv, a.idesc, (kb | ks) != 0);
              tc_mma_f16(d_tmem, a_hi + adv, b_lo + adv, a.idesc, 1);
              tc_mma_f16(d_tmem, a_lo + adv, b_hi + adv, a.idesc, 1);
            }
            tc_commit_mask(empty_bar + 8 * stage, pair_mask);
            if (kb == nb - 1) tc_commit_mask(tfull_bar + 8 * acc, pair_mask);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================================================== epilogue warps: TMEM -> smem -> TMA store
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t stg = epi_base + (warp - 2) * EPI_WARP_BYTES;
    const uint32_t sw = (uint32_t)(lane & 7);
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int b = tile / tiles_per_chunk, rem = tile - b * tiles_per_chunk;
      const int tm = rem / a.n_tn, tn = rem - tm * a.n_tn;
      const int mrow = tm * 256 + (int)rank * 128 + q * 32;            // first output row of this warp (M % 32 == 0)
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      if (mrow < a.M) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + half * 128);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int col0 = tn * 256 + half * 128 + c * 32;
          if (col0 >= a.Nn) break;
          float v[32];
          tc_ld_32x32(taddr + c * 32, v);
          tc_ld_wait();
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          const uint32_t sb = stg + (uint32_t)lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((((uint32_t)j) ^ sw) << 4)), "f"(v[4 * j]),
                         "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmP, stg, col0, b * a.M + mrow);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa(tempty_bar + 8 * acc, 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// dW[t] tile (64 x 64) = sum over the type's chunk partials + the tail rows [tail0, end) of the type, operands
// recombined from their [hi; lo] planes.  256 threads, 4 x 4 outputs each.
struct WgradTypes {
  int T;
  int chunk_ptr[WSI_MAX_TYPES + 1];
  int tail0[WSI_MAX_TYPES];
  int end[WSI_MAX_TYPES];
};

__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    int n_rows, int M, int Nn, const __grid_constant__ WgradTypes ty, float* __restrict__ dw) {
  __shared__ float As[64][64 + 4];      // [row][m]
  __shared__ float Bs[64][64 + 4];      // [row][n]
  wsi_pdl_wait();
  const int t = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, tyy = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nc = n0 + tx * 4;
  const int64_t plane = (int64_t)M * Nn;
  if (nc < Nn) {
    for (int b = ty.chunk_ptr[t]; b < ty.chunk_ptr[t + 1]; ++b) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + tyy * 4 + i;
        if (m < M) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(partial + b * plane + (int64_t)m * Nn + nc));
          acc[i][0] += p.x; acc[i][1] += p.y; acc[i][2] += p.z; acc[i][3] += p.w;
        }
      }
    }
  }
  const int r0 = ty.tail0[t], nr = ty.end[t] - r0;                  // nr < 64
  if (nr > 0) {
    for (int i = tid; i < nr * 64; i += 256) {
      const int r = i >> 6, c = i & 63;
      const int64_t row = r0 + r;
      float av = 0.f, bv = 0.f;
      if (m0 + c < M) av = __bfloat162float(dy[row * M + m0 + c]) + __bfloat162float(dy[(row + n_rows) * M + m0 + c]);
      if (n0 + c < Nn) bv = __bfloat162float(x[row * Nn + n0 + c]) + __bfloat162float(x[(row + n_rows) * Nn + n0 + c]);
      As[r][c] = av;
      Bs[r][c] = bv;
    }
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[r][tyy * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[r][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  if (nc < Nn) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + tyy * 4 + i;
      if (m < M)
        *reinterpret_cast<float4*>(dw + t * plane + (int64_t)m * Nn + nc) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// Cuts the whole 64-row blocks of every type into chunks: ~2 (chunk, tile) work items per CTA pair.
void plan_chunks(const int32_t* tp, int T, int tiles_per_chunk, int sms, WgradChunks* ch, WgradTypes* ty) {
  int64_t total_blocks = 0;
  for (int t = 0; t < T; ++t) total_blocks += (tp[t + 1] - tp[t]) / BK;
  int target = (2 * (sms / 2) + tiles_per_chunk - 1) / tiles_per_chunk;
  if (target < 1) target = 1;
  if (target > MAX_CHUNKS - T) target = MAX_CHUNKS - T;
  int64_t len = (total_blocks + target - 1) / target;
  if (len < 4) len = 4;                                              // a chunk shorter than the ring is all prologue
  ch->n = 0;
  ty->T = T;
  for (int t = 0; t < T; ++t) {
    ty->chunk_ptr[t] = ch->n;
    const int blocks = (tp[t + 1] - tp[t]) / BK;
    const int nck = (int)((blocks + len - 1) / len);
    for (int c = 0; c < nck; ++c) {
      const int b0 = (int)((int64_t)blocks * c / nck), b1 = (int)((int64_t)blocks * (c + 1) / nck);
      ch->row0[ch->n] = tp[t] + b0 * BK;
      ch->nblk[ch->n] = b1 - b0;
      ++ch->n;
    }
    ty->tail0[t] = tp[t] + blocks * BK;
    ty->end[t] = tp[t + 1];
  }
  ty->chunk_ptr[T] = ch->n;
}

inline int tiles_of(int M, int Nn) { return ((M + 255) / 256) * ((Nn + 255) / 256); }

}  // namespace

extern "C" int wsi_typed_wgrad_supported(int64_t n_rows, int M, int Nn, int T) {
  return n_rows >= 512 && n_rows < (1ll << 30) && M >= 64 && M % 32 == 0 && Nn >= 64 && Nn % 8 == 0 && T >= 1 &&
         T <= WSI_MAX_TYPES && MAX_CHUNKS - T >= 1 && (int64_t)MAX_CHUNKS * M < (1ll << 31);
}

extern "C" int64_t wsi_typed_wgrad_workspace_bytes(int M, int Nn, const int32_t* type_ptr_host, int T) {
  if (!type_ptr_host || T < 1 || T > WSI_MAX_TYPES) return -1;
  int sms = wsi_num_sms();
  if (sms <= 0) return -1;
  static thread_local WgradChunks ch;
  static thread_local WgradTypes ty;
  plan_chunks(type_ptr_host, T, tiles_of(M, Nn), sms, &ch, &ty);
  return (int64_t)(ch.n > 0 ? ch.n : 1) * M * Nn * 4 + 256;
}

extern "C" int wsi_typed_wgrad(const void* dy_op, const void* x_op, int M, int Nn, const int32_t* type_ptr_host, int T,
                               float* dw, void* workspace, int64_t workspace_bytes, void* stream_) {
  WSI_CHECK_ARG(dy_op && x_op && dw && type_ptr_host, "typed_wgrad: null pointer");
  WSI_CHECK_ARG(T >= 1 && T <= WSI_MAX_TYPES, "typed_wgrad: T=%d out of range", T);
  const int64_t n_rows = type_ptr_host[T];
  WSI_CHECK_ARG(wsi_typed_wgrad_supported(n_rows, M, Nn, T), "typed_wgrad: shape N=%lld M=%d Nn=%d not supported",
                (long long)n_rows, M, Nn);
  WSI_CHECK_ARG((reinterpret_cast<uintptr_t>(dy_op) & 127) == 0 && (reinterpret_cast<uintptr_t>(x_op) & 127) == 0 &&
                    (reinterpret_cast<uintptr_t>(dw) & 15) == 0,
                "typed_wgrad: operands must be 128 B aligned, dw 16 B aligned");
  cudaStream_t stream = wsi_stream(stream_);
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  WgradChunks ch;
  WgradTypes ty;
  WgradArgs a{};
  a.n_rows = (int)n_rows; a.M = M; a.Nn = Nn;
  a.n_tm = (M + 255) / 256; a.n_tn = (Nn + 255) / 256;
  plan_chunks(type_ptr_host, T, a.n_tm * a.n_tn, sms, &ch, &ty);
  const int64_t need = (int64_t)(ch.n > 0 ? ch.n : 1) * M * Nn * 4 + 256;
  WSI_CHECK_ARG(workspace && workspace_bytes >= need, "typed_wgrad: workspace of %lld bytes needed", (long long)need);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  if (ch.n > 0) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
      attr_err = cudaFuncSetAttribute((const void*)typed_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    });
    WSI_CHECK_CUDA(attr_err);
    CUtensorMap tmA, tmB, tmP;
    int rc = make_map(&tmA, dy_op, 2 * n_rows, M, (int64_t)M * 2, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    if (rc != WSI_OK) return rc;
    rc = make_map(&tmB, x_op, 2 * n_rows, Nn, (int64_t)Nn * 2, 64, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
    if (rc != WSI_OK) return rc;
    rc = make_map(&tmP, partial, (int64_t)ch.n * M, Nn, (int64_t)Nn * 4, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc != WSI_OK) return rc;
    // kind::f16, bf16 x bf16 -> fp32, A and B MN-major (bits 15 / 16), N = 256, M = 256
    a.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const int total = ch.n * a.n_tm * a.n_tn;
    const int max_clusters = sms / 2;
    const int clusters = total < max_clusters ? total : max_clusters;
    cudaError_t le = wsi_launch_pdl(typed_wgrad_tc_kernel, dim3(2 * clusters), dim3(THREADS), SMEM_BYTES, stream, tmA, tmB,
                                    tmP, ch, a);
    WSI_CHECK_CUDA(le);
    WSI_CHECK_LAUNCH();
  }
  const dim3 rgrid((Nn + 63) / 64, (M + 63) / 64, T);
  cudaError_t le = wsi_launch_pdl(wgrad_reduce_kernel, rgrid, dim3(256), 0, stream, (const float*)partial,
                                  (const __nv_bfloat16*)dy_op, (const __nv_bfloat16*)x_op, (int)n_rows, M, Nn, ty, dw);
  WSI_CHECK_CUDA(le);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
