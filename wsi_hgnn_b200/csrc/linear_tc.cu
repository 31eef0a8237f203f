// Typed (grouped by node type) linear on the 5th-gen tensor cores: tcgen05.mma, TMEM accumulators, TMA operands.
//
//   Y[rows of type t] = epilogue( X[rows of type t] . W[t]^T )                       (kernel K1)
// replaces the per-node-type nn.Linear calls of models/HEATNet4.py:100-102,134,202 / models/HGT.py:82-84,121,180.
//
// fp32 parity (1e-3 relative on the logits after L layers) rules out a single TF32 pass, so every fp32 operand
// is split into two bf16 values  x = hi + lo  (|lo| <= 2^-9 |x|)  and the product is formed from three
// bf16 x bf16 -> fp32 MMAs   hi.hi + hi.lo + lo.hi   (the dropped lo.lo term is ~2^-18 relative):
//   1. split_bf16_kernel      X, W (fp32) -> workspace [X_hi; X_lo] and [W_hi; W_lo]   (HBM-bound pre-pass)
//   2. typed_linear_tc_kernel persistent, warp-specialised, one CTA per SM:
//        warp 0      TMA producer: per k-block loads the four 128B-swizzled tiles A_hi, A_lo, B_hi, B_lo
//        warp 1      MMA issuer  : 3 tcgen05.mma (M=128, N=256, K=16) per 16-wide k-slice into TMEM
//        warps 2..5  epilogue    : tcgen05.ld -> smem transpose -> fused epilogue (epilogue.cuh) -> coalesced
//                                  16 B global stores; double-buffered TMEM (2 x 256 columns) so the epilogue of
//                                  tile i overlaps the MMAs of tile i+1
// Tensor-pipe bound: algorithmic flops 2*N*K*n_out (x3 MMAs issued for the split).
#include <cuda.h>

#include <mutex>

#include "epilogue.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 2, UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int B_TILE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
constexpr int EPI_LD = 36;                               // staging row pitch in floats (16 B aligned, conflict free)
constexpr int EPI_WARP_FLOATS = 32 * EPI_LD;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 4 * EPI_WARP_FLOATS * 4 + 256;
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 512;

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row swizzle atoms of 1024 B):
// start address >> 4 | LBO 1 (unused for swizzled K-major) | SBO 1024 B >> 4 | version 1 (sm_100) | SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, A = B = bf16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ------------------------------------------------------------------------------------------ pre-pass
// fp32 -> [hi; lo] bf16.  src rows have stride ld_src floats; dst is dense [2 * rows, K] (lo half at row `rows`).
struct SplitJob { const float* src; int64_t ld_src; int64_t rows; __nv_bfloat16* dst; };

__global__ void __launch_bounds__(256) split_bf16_kernel(SplitJob a, SplitJob b, int K) {
  const int kv = K >> 2;                                  // float4 groups per row (K % 8 == 0)
  const int64_t na = a.rows * kv, total = na + b.rows * kv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const SplitJob& j = i < na ? a : b;
    const int64_t li = i < na ? i : i - na;
    const int64_t r = li / kv;
    const int c = (int)(li - r * kv) << 2;
    const float4 x = __ldg(reinterpret_cast<const float4*>(j.src + r * j.ld_src + c));
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    wsi_split_bf16(x.x, h0, l0); wsi_split_bf16(x.y, h1, l1); wsi_split_bf16(x.z, h2, l2); wsi_split_bf16(x.w, h3, l3);
    __nv_bfloat162 hv[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
    __nv_bfloat162 lv[2] = {__halves2bfloat162(l0, l1), __halves2bfloat162(l2, l3)};
    *reinterpret_cast<uint2*>(j.dst + r * K + c) = *reinterpret_cast<uint2*>(hv);
    *reinterpret_cast<uint2*>(j.dst + (j.rows + r) * K + c) = *reinterpret_cast<uint2*>(lv);
  }
}

// ------------------------------------------------------------------------------------------ main kernel
struct TcArgs {
  int n_rows;        // N  (row offset of the lo half of the A workspace)
  int w_rows;        // T * n_out (row offset of the lo half of the W workspace)
  int K, n_out;
  int n_tiles_m, n_tiles_n;
  int vec_epi;       // 1: every epilogue operand is 16 B aligned -> float4 loads
};

__device__ __forceinline__ float4 epi_apply4(const LinearEpilogue& ep, float4 acc, int t, int64_t row, int n,
                                             float alpha, bool gate_open, float rscale, bool vec) {
  float a[4] = {acc.x, acc.y, acc.z, acc.w};
  if (vec) {
    float bb[4] = {0.f, 0.f, 0.f, 0.f}, mm[4] = {1.f, 1.f, 1.f, 1.f}, rr[4] = {0.f, 0.f, 0.f, 0.f};
    if (ep.bias) *reinterpret_cast<float4*>(bb) = __ldg(reinterpret_cast<const float4*>(ep.bias + (int64_t)t * ep.n_out + n));
    if (ep.drop_mask) *reinterpret_cast<float4*>(mm) = __ldg(reinterpret_cast<const float4*>(ep.drop_mask + row * ep.ldmask + n));
    if (ep.skip) *reinterpret_cast<float4*>(rr) = __ldg(reinterpret_cast<const float4*>(ep.res + row * ep.ldres + n));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = a[i] + bb[i];
      if (ep.act == WSI_ACT_GELU) v = wsi_gelu(v);
      v *= mm[i];
      if (ep.skip) v = gate_open ? (v * alpha + rr[i] * (1.0f - alpha)) : rr[i];
      a[i] = v * rscale;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = wsi_epilogue_value(ep, a[i], t, row, n + i, alpha, gate_open, rscale);
  }
  return make_float4(a[0], a[1], a[2], a[3]);
}

__global__ void __launch_bounds__(THREADS, 1)
typed_linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ TypeSegs segs, const __grid_constant__ LinearEpilogue ep, TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                      // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen = smem_raw + (base - raw);
  float* epi_stage = reinterpret_cast<float*>(gen + STAGES * STAGE_BYTES);
  const uint32_t bars = base + STAGES * STAGE_BYTES + 4 * EPI_WARP_FLOATS * 4;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES,
                 tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (a.K + BK - 1) / BK;
  const int total_tiles = a.n_tiles_m * a.n_tiles_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tm = tile / a.n_tiles_n, tn = tile - tm * a.n_tiles_n;
        const int t = wsi_tile_group(segs, tm);
        const int row0 = segs.ptr[t] + (tm - segs.tile_start[t]) * BM;
        const int wrow0 = t * a.n_out + tn * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          const uint32_t s0 = base + stage * STAGE_BYTES;
          mbar_expect_tx(fb, STAGE_BYTES);
          tma_load_2d(&tmA, fb, s0, kb * BK, row0);
          tma_load_2d(&tmA, fb, s0 + A_TILE_BYTES, kb * BK, a.n_rows + row0);
          tma_load_2d(&tmB, fb, s0 + 2 * A_TILE_BYTES, kb * BK, wrow0);
          tma_load_2d(&tmB, fb, s0 + 2 * A_TILE_BYTES + B_TILE_BYTES, kb * BK, a.w_rows + wrow0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one elected lane)
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);                // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t s0 = base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc(s0), a_lo = make_smem_desc(s0 + A_TILE_BYTES);
          const uint64_t b_hi = make_smem_desc(s0 + 2 * A_TILE_BYTES),
                         b_lo = make_smem_desc(s0 + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);   // +32 B per k-slice inside the swizzle atom
            tc_mma_bf16(d_tmem, a_hi + adv, b_hi + adv, IDESC, (kb | ks) != 0);
            tc_mma_bf16(d_tmem, a_hi + adv, b_lo + adv, IDESC, 1);
            tc_mma_bf16(d_tmem, a_lo + adv, b_hi + adv, IDESC, 1);
          }
          tc_commit(empty_bar + 8 * stage);                          // smem slot free once these MMAs retire
          if (kb == num_kb - 1) tc_commit(tfull_bar + 8 * acc);      // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue warps (TMEM lane quarter = warp % 4)
    const int q = warp & 3;
    float* stg = epi_stage + (warp - 2) * EPI_WARP_FLOATS;
    const bool vec = a.vec_epi != 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tm = tile / a.n_tiles_n, tn = tile - tm * a.n_tiles_n;
      const int t = wsi_tile_group(segs, tm);
      const int row0 = segs.ptr[t] + (tm - segs.tile_start[t]) * BM + q * 32;
      const int row_end = segs.ptr[t + 1];
      const int n0 = tn * BN;
      const float alpha = ep.skip ? wsi_sigmoid(__ldg(ep.skip + t)) : 1.0f;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * BN;
      const int rsub = lane >> 3, c4 = (lane & 7) << 2;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (n0 + c * 32 >= a.n_out || row0 >= row_end) break;        // warp-uniform
        float v[32];
        tc_ld_32x32(taddr + c * 32, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * EPI_LD + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int n = n0 + c * 32 + c4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsub;
          const int64_t row = row0 + rr;
          if (row < row_end && n < a.n_out) {
            const float4 accv = *reinterpret_cast<const float4*>(stg + rr * EPI_LD + c4);
            const bool gate_open = ep.row_gate ? __ldg(ep.row_gate + row) != 0.f : true;
            const float rscale = ep.row_scale ? __ldg(ep.row_scale + row) : 1.0f;
            const float4 o = epi_apply4(ep, accv, t, row, n, alpha, gate_open, rscale, vec);
            *reinterpret_cast<float4*>(ep.y + row * ep.ldy + n) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 [rows, K] row-major, box [box_rows, BK], 128 B swizzle, out-of-bounds elements read as 0
int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { wsi_set_error("typed_linear(tcgen05): cuTensorMapEncodeTiled is not available"); return WSI_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { wsi_set_error("typed_linear(tcgen05): cuTensorMapEncodeTiled failed (%d)", (int)r); return WSI_ERR_CUDA; }
  return WSI_OK;
}

inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Shapes the tensor-core path takes: K a multiple of 8 (16 B rows for TMA / float4 split), output rows 16 B
// aligned, and enough work to fill 128-row MMA tiles (small [B, D] readout GEMMs stay on the SIMT path).
bool wsi_typed_linear_tc_supported(int64_t n_rows, int K, int n_out, int64_t ldx) {
  return n_rows >= 512 && n_rows < (1ll << 30) && K >= 64 && K % 8 == 0 && n_out >= 64 && n_out % 4 == 0 &&
         ldx % 4 == 0 && (int64_t)n_out * WSI_MAX_TYPES < (1ll << 30);
}

int64_t wsi_typed_linear_tc_workspace(int64_t n_rows, int K, int n_out, int T) {
  return align256(2 * n_rows * K * 2) + align256(2 * (int64_t)T * n_out * K * 2) + 1024;
}

int wsi_typed_linear_tc_launch(const float* x, int64_t ldx, const float* w, int K, const int32_t* type_ptr_host,
                               int T, const LinearEpilogue& ep, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream) {
  const int64_t n_rows = type_ptr_host[T];
  const int n_out = ep.n_out;
  WSI_CHECK_ARG(workspace && workspace_bytes >= wsi_typed_linear_tc_workspace(n_rows, K, n_out, T),
                "typed_linear(tcgen05): workspace of %lld bytes needed", (long long)wsi_typed_linear_tc_workspace(n_rows, K, n_out, T));
  WSI_CHECK_ARG(aligned16(x) && aligned16(ep.y) && ep.ldy % 4 == 0,
                "typed_linear(tcgen05): x / y must be 16 B aligned with row strides that are multiples of 4 floats");
  uintptr_t wsp = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023;
  __nv_bfloat16* a_ws = reinterpret_cast<__nv_bfloat16*>(wsp);
  __nv_bfloat16* w_ws = reinterpret_cast<__nv_bfloat16*>(wsp + align256(2 * n_rows * K * 2));

  TypeSegs segs;
  if (wsi_make_segs(&segs, type_ptr_host, T, BM) != 0) { wsi_set_error("typed_linear: bad type_ptr"); return WSI_ERR_ARG; }

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(typed_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  WSI_CHECK_CUDA(attr_err);
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;

  // 1. split pre-pass
  SplitJob ja{x, ldx, n_rows, a_ws}, jb{w, (int64_t)K, (int64_t)T * n_out, w_ws};
  const int64_t groups = (ja.rows + jb.rows) * (K / 4);
  int sblocks = (int)((groups + 255) / 256);
  if (sblocks > sms * 8) sblocks = sms * 8;
  split_bf16_kernel<<<sblocks, 256, 0, stream>>>(ja, jb, K);
  WSI_CHECK_LAUNCH();

  // 2. tensor-core GEMM
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, a_ws, 2 * n_rows, K, BM);
  if (rc != WSI_OK) return rc;
  rc = make_map(&tmB, w_ws, 2 * (int64_t)T * n_out, K, BN);
  if (rc != WSI_OK) return rc;
  TcArgs a{};
  a.n_rows = (int)n_rows; a.w_rows = T * n_out; a.K = K; a.n_out = n_out;
  a.n_tiles_m = segs.tile_start[T]; a.n_tiles_n = (n_out + BN - 1) / BN;
  a.vec_epi = (!ep.bias || (aligned16(ep.bias) && n_out % 4 == 0)) &&
              (!ep.drop_mask || (aligned16(ep.drop_mask) && ep.ldmask % 4 == 0)) &&
              (!ep.res || (aligned16(ep.res) && ep.ldres % 4 == 0));
  const int total = a.n_tiles_m * a.n_tiles_n;
  if (total == 0) return WSI_OK;
  const int grid = total < sms ? total : sms;
  typed_linear_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmB, segs, ep, a);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
