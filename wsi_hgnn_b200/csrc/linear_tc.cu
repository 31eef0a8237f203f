// Typed (grouped by node type) linear on the 5th-gen tensor cores: tcgen05.mma, TMEM accumulators, TMA operands.
//
//   Y[rows of type t] = epilogue( X[rows of type t] . W[t]^T )                       (kernel K1)
// replaces the per-node-type nn.Linear calls of models/HEATNet4.py:100-102,134,202 / models/HGT.py:82-84,121,180.
//
// fp32 parity (1e-3 relative on the logits after L layers) rules out a single TF32 pass, so every fp32 operand
// is split into two bf16 values  x = hi + lo  (|lo| <= 2^-9 |x|)  and the product is formed from three
// bf16 x bf16 -> fp32 MMAs   hi.hi + hi.lo + lo.hi   (the dropped lo.lo term is ~2^-18 relative):
//   1. split_bf16_kernel      X, W (fp32) -> workspace [X_hi; X_lo] and [W_hi; W_lo]   (HBM-bound pre-pass)
//   2. typed_linear_tc_kernel persistent, warp-specialised, one CTA PAIR (cluster of 2, cta_group::2) per two SMs;
//      a pair owns a 256 x 256 output tile: each CTA stages its own 128 rows of A and HALF of the W tile, so the
//      L2 -> smem traffic per output (the limiter of the 3-term product) is 2/3 of a single-CTA 128 x 256 tile:
//        warp 0      TMA producer (each CTA): per k-block the four 128B-swizzled tiles A_hi, A_lo, B_hi/2, B_lo/2,
//                    completion bytes posted on the LEADER CTA's full barrier
//        warp 1      MMA issuer (leader CTA): 3 tcgen05.mma.cta_group::2 (M=256, N=256, K=16) per 16-wide k-slice,
//                    tcgen05.commit multicast to both CTAs' empty / tmem-full barriers
//        warps 2..5  epilogue (each CTA, its own 128 TMEM lanes): tcgen05.ld -> smem transpose -> fused epilogue
//                    (epilogue.cuh) -> coalesced 16 B global stores; double-buffered TMEM (2 x 256 columns) so the
//                    epilogue of tile i overlaps the MMAs of tile i+1
// Tensor-pipe bound: algorithmic flops 2*N*K*n_out (x3 MMAs issued for the split).
//
// Operand formats (WSI_OPF_*, include/wsi_hgnn.h).  The kernel is a template on TERMS:
//   TERMS 3  WSI_OPF_BF16X3: the [hi; lo] scheme above (~2^-17 relative: "exact" mode, gradients)
//   TERMS 1  WSI_OPF_F16 / WSI_OPF_BF16: ONE 16-bit operand per matrix, one MMA per k-slice, a 6-stage ring of
//            32 KB stages.  fp16 (11-bit significand = TF32's) is the default of the fp32 models: measured on the 16
//            reference-generated goldens + config 2 (tools/precision_study.py, profiles/r2_precision_study.json) the
//            whole forward stays 3.6x inside the 1e-3 parity bar, at 1/3 of the tensor work and 1/2 of the L2 -> smem
//            operand bytes of the 3-term product; bf16 serves the bf16-storage configuration (BASELINE config 3).
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "epilogue.cuh"
#include "operand.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;            // rows per CTA (TMEM lanes)
constexpr int BN = 256;            // columns per tile = UMMA N; each CTA of the pair stages BN/2 rows of W
constexpr int BK = 64;             // 16-bit elements per k-block = one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int PAIR_M = 2 * BM;     // rows per CTA pair (UMMA M = 256, cta_group::2)
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int B_TILE_BYTES = (BN / 2) * BK * 2;
constexpr int EPI_LD = 32;                               // staging row pitch in floats; float4 slots XOR-swizzled by row
constexpr int EPI_WARP_FLOATS = 32 * EPI_LD;
constexpr int EPI_WARPS = 8;                             // 2 warps per TMEM lane quarter, each takes half of the columns
// Tile shapes that were built, verified bit-identical and measured in round 1 and do NOT ship any more (they lost):
// 128-wide tiles with a 4-stage ring (more L2 -> smem operand bytes per flop) and a 4-CTA cluster sharing the A operand
// by TMA multicast (no gain: only 132 of 148 SMs host 4-CTA clusters).  See DESIGN.md section 3.
constexpr int stages_of(int terms) { return terms == 3 ? 3 : 5; }
constexpr int stage_bytes_of(int terms) { return (terms == 3 ? 2 : 1) * (A_TILE_BYTES + B_TILE_BYTES); }       // per CTA
// epilogue staging: 4 KB (32 rows x 128 B) buffers per warp; the single-operand kernel has room for two (the TMA store
// of chunk c drains while chunk c+1 is staged; the second one doubles as the 16-bit staging when y_op is written)
constexpr int epi_bufs_of(int terms) { return terms == 3 ? 1 : 2; }
constexpr int smem_bytes_of(int terms) { return 1024 + stages_of(terms) * stage_bytes_of(terms) + EPI_WARPS * epi_bufs_of(terms) * EPI_WARP_FLOATS * 4 + 256; }
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr uint32_t TMEM_COLS = 512;

// kind::f16 instruction descriptor: fp32 accumulate (bit 4), A / B format at bits 7 / 10 (0 = fp16, 1 = bf16), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24
inline uint32_t idesc_of(bool bf16) {
  const uint32_t f = bf16 ? 1u : 0u;
  return (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PAIR_M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ pre-pass
// fp32 -> operand form.  src rows have stride ld_src floats; dst is dense:
//   WSI_OPF_BF16X3  [2 * rows, K] bf16 (hi rows, then lo rows at row `rows`)
//   WSI_OPF_F16     [rows, K] fp16, round to nearest, clamped to the finite fp16 range (+-65504)
//   WSI_OPF_BF16    [rows, K] bf16
struct SplitJob { const float* src; int64_t ld_src; int64_t rows; void* dst; const int* row_idx; };   // row_idx: optional gather of the source rows

template <int OPF>
__global__ void __launch_bounds__(256) convert_operand_kernel(SplitJob a, SplitJob b, int K) {
  wsi_pdl_trigger();                                      // the GEMM that follows may set itself up while this drains
  const int kv = K >> 2;                                  // float4 groups per row (K % 8 == 0)
  const int64_t na = a.rows * kv, total = na + b.rows * kv;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // four independent 16 B loads in flight per thread (HBM-bound: 4 B in, 2-4 B out per element)
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 x[4];
    int64_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = i0 + u * stride;
      if (idx[u] < total) {
        const SplitJob& j = idx[u] < na ? a : b;
        const int64_t li = idx[u] < na ? idx[u] : idx[u] - na;
        const int64_t r = li / kv;
        const int c = (int)(li - r * kv) << 2;
        const int64_t sr = j.row_idx ? (int64_t)__ldg(j.row_idx + r) : r;
        x[u] = __ldg(reinterpret_cast<const float4*>(j.src + sr * j.ld_src + c));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (idx[u] < total) {
        const SplitJob& j = idx[u] < na ? a : b;
        const int64_t li = idx[u] < na ? idx[u] : idx[u] - na;
        const int64_t r = li / kv;
        const int c = (int)(li - r * kv) << 2;
        store_operand4<OPF>(reinterpret_cast<uint16_t*>(j.dst) + r * K + c, j.rows * K, x[u]);
      }
    }
  }
}

// fp32 -> [hi; lo] operand form AND the per-type column sums of the same matrix in one pass: the backward of a typed
// linear needs dY as a GEMM operand (data / weight gradient) and sum_rows dY as the bias gradient - two full reads of dY
// otherwise.  One block per tile of CS_ROWS rows of ONE type (tiles never straddle types); a thread owns 4 columns, walks
// the rows of the tile (4 loads in flight) and leaves one partial sum per tile; colsum_finish_kernel adds the partials
// of a type in a fixed order (deterministic, no atomics).
constexpr int CS_ROWS = 256;

__global__ void __launch_bounds__(256) convert_colsum_kernel(const float* __restrict__ src, int64_t ld, int64_t n_rows, int K,
                                                             const __grid_constant__ TypeSegs segs, uint16_t* __restrict__ dst,
                                                             float* __restrict__ partial) {
  const int tile = blockIdx.x;
  const int t = wsi_tile_group(segs, tile);
  const int r0 = segs.ptr[t] + (tile - segs.tile_start[t]) * CS_ROWS;
  const int r1 = min(r0 + CS_ROWS, segs.ptr[t + 1]);
  const int64_t lo_off = n_rows * K;
  for (int c = threadIdx.x * 4; c < K; c += 1024) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0; r < r1; r += 4) {
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + u < r1) x[u] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(r + u) * ld + c));
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r + u < r1) {
          store_operand4<WSI_OPF_BF16X3>(dst + (int64_t)(r + u) * K + c, lo_off, x[u]);
          acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
        }
    }
    *reinterpret_cast<float4*>(partial + (int64_t)tile * K + c) = acc;
  }
}

// out[t, c] = sum over the tiles of type t of partial[tile, c]; block = 64 columns x 4 tile lanes
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int K,
                                                            const __grid_constant__ TypeSegs segs, float* __restrict__ out) {
  __shared__ float red[4][64];
  const int t = blockIdx.y;
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), q = threadIdx.x >> 6;
  float acc = 0.f;
  if (c < K)
    for (int tile = segs.tile_start[t] + q; tile < segs.tile_start[t + 1]; tile += 4) acc += __ldg(partial + (int64_t)tile * K + c);
  red[q][threadIdx.x & 63] = acc;
  __syncthreads();
  if (q == 0 && c < K) out[(int64_t)t * K + c] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

// dst row i = src row row_idx[i] of a dense 16-bit [*, K] matrix, 16 bytes (8 elements) per thread
__global__ void __launch_bounds__(256) gather_rows16_kernel(const uint4* __restrict__ src, int64_t ld4, const int* __restrict__ row_idx,
                                                            int64_t rows, int kv, uint4* __restrict__ dst) {
  const int64_t total = rows * kv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / kv;
    const int c = (int)(i - r * kv);
    dst[i] = __ldg(src + (int64_t)__ldg(row_idx + r) * ld4 + c);
  }
}

// ------------------------------------------------------------------------------------------ main kernel
struct TcArgs {
  int n_rows;        // N  (TERMS 3: row offset of the lo half of the A workspace)
  int w_rows;        // T * n_out (TERMS 3: row offset of the lo half of the W workspace)
  int K, n_out;
  int n_tiles_m, n_tiles_n;   // n_tiles_m counts 256-row PAIR tiles
  void* y_op;        // optional operand-form copy of y (the next GEMM's A operand): TERMS 3 bf16 [2 * n_rows, n_out]
                     // (hi; lo), TERMS 1 fp16 / bf16 [n_rows, n_out]
  uint32_t idesc;
  int tma_store;     // y (and, TERMS 1, y_op) described by tmY / tmY16: the plain epilogue may use TMA stores
  int op_bf16;       // TERMS 1: 16-bit type of the operands and of y_op (0 = fp16, 1 = bf16)
  int dbg;           // development only (wsi_dev_set("tc_debug")): bit 0 = skip the MMAs, bit 1 = skip the TMA loads,
                     // bit 2 = skip the epilogue body, bit 3 = no global stores in the epilogue, bit 4 = no TMEM reads
};

// FULL = false: v = act(acc + bias).  FULL = true: + dropout mask, sigma(skip) residual mix with row gate, row scale.
template <bool FULL, bool GELU>
__device__ __forceinline__ float4 epi_mix4(const LinearEpilogue& ep, float4 acc, float4 bb, float4 mm, float4 rr,
                                           float alpha, bool gate_open, float rscale) {
  float a[4] = {acc.x, acc.y, acc.z, acc.w};
  const float b4[4] = {bb.x, bb.y, bb.z, bb.w}, m4[4] = {mm.x, mm.y, mm.z, mm.w}, r4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = a[i] + b4[i];
    if (GELU) v = wsi_gelu(v);
    if (FULL) {
      v *= m4[i];
      if (ep.skip) v = gate_open ? (v * alpha + r4[i] * (1.0f - alpha)) : r4[i];
      v *= rscale;
    }
    a[i] = v;
  }
  return make_float4(a[0], a[1], a[2], a[3]);
}

template <bool FULL, bool GELU, int TERMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
typed_linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmY16,
                       const __grid_constant__ TypeSegs segs, const __grid_constant__ LinearEpilogue ep, TcArgs a) {
  constexpr int STAGES = stages_of(TERMS), STAGE_BYTES = stage_bytes_of(TERMS), EPI_BUFS = epi_bufs_of(TERMS);
  constexpr int B_OFF = (TERMS == 3 ? 2 : 1) * A_TILE_BYTES;          // first W tile inside a stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                      // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* gen = smem_raw + (base - raw);
  float* epi_stage = reinterpret_cast<float*>(gen + STAGES * STAGE_BYTES);
  const uint32_t bars = base + STAGES * STAGE_BYTES + EPI_WARPS * EPI_BUFS * EPI_WARP_FLOATS * 4;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES, tfull_bar = bars + 16 * STAGES,
                 tempty_bar = tfull_bar + 16, tmem_slot = tempty_bar + 16;
  volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                           // 0 = leader (issues the MMAs of the pair)
  const uint16_t pair_mask = 3;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int num_kb = (a.K + BK - 1) / BK;
  const int n_tn = a.n_tiles_n;
  const int total_tiles = a.n_tiles_m * n_tn;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (a.tma_store) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
      if (a.y_op) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY16) : "memory");
    }
    // full (the leader's is the one used): ONE arrival, the leader's expect_tx for the bytes of BOTH CTAs - the peer's
    // loads post their bytes on it and need no arrival of their own (a phase completes only when the pending arrival
    // AND the byte count reach zero, whichever CTA is ahead).  The peer used to add a release.cluster arrive per
    // k-block: a MEMBAR + ERRBAR that drained its just-issued bulk loads, i.e. one full L2 round trip per k-block
    // (measured round 2: "loads only" 28.7 us for 99 MB and 32.8 us for 198 MB - latency, not bytes).
    // empty / tmem_full: one tcgen05.commit; tmem_empty (leader's is the one waited on): the epilogue warps of both CTAs
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
  cluster_sync();                                                    // peer barriers are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  // everything above (barriers, TMEM, tensor-map prefetch) may run while the previous kernel of the stream drains;
  // from here on its outputs are read
  wsi_pdl_trigger();
  wsi_pdl_wait();

  if (warp == 0) {
    // ===================================================================== TMA producer (one lane per CTA)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        const int tm = tile / n_tn, tn = tile - tm * n_tn;
        const int t = wsi_tile_group(segs, tm);
        const int row0 = segs.ptr[t] + (tm - segs.tile_start[t]) * PAIR_M + (int)rank * BM;
        const int wrow0 = t * a.n_out + tn * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = mapa(full_bar + 8 * stage, 0);
          const uint32_t s0 = base + stage * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar + 8 * stage, (a.dbg & 2) ? 0 : 2 * STAGE_BYTES);
          if (!(a.dbg & 2)) {
            tma_load_2d(&tmA, fb, s0, kb * BK, row0);
            if (TERMS == 3) tma_load_2d(&tmA, fb, s0 + A_TILE_BYTES, kb * BK, a.n_rows + row0);
            tma_load_2d(&tmB, fb, s0 + B_OFF, kb * BK, wrow0);
            if (TERMS == 3) tma_load_2d(&tmB, fb, s0 + B_OFF + B_TILE_BYTES, kb * BK, a.w_rows + wrow0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA, one elected lane)
    if (rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = a.idesc;
      for (int tile = pair; tile < total_tiles; tile += n_pairs) {
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);              // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t s0 = base + stage * STAGE_BYTES;
            const uint64_t a_hi = make_smem_desc(s0), b_hi = make_smem_desc(s0 + B_OFF);
            if (!(a.dbg & 1))
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);   // +32 B per k-slice inside the swizzle atom
              tc_mma_f16(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | ks) != 0);
              if (TERMS == 3) {
                const uint64_t a_lo = make_smem_desc(s0 + A_TILE_BYTES), b_lo = make_smem_desc(s0 + B_OFF + B_TILE_BYTES);
                tc_mma_f16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                tc_mma_f16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1);
              }
            }
            tc_commit_mask(empty_bar + 8 * stage, pair_mask);        // smem slot free once these retire (both CTAs)
            if (kb == num_kb - 1) tc_commit_mask(tfull_bar + 8 * acc, pair_mask);   // accumulator complete (both CTAs)
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================================================== epilogue warps
    // TMEM lane quarter = warp % 4 (hardware rule); the two warps of a quarter split the BN columns in halves.
    const int q = warp & 3, half = (warp - 2) >> 2;
    float* stg = epi_stage + (warp - 2) * EPI_BUFS * EPI_WARP_FLOATS;
    const uint32_t stg_u32 = smem_u32(stg);
    uint32_t store_buf = 0;                                          // staging buffer of the next TMA store (EPI_BUFS == 2)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    const int rsub = lane >> 3, c4 = (lane & 7) << 2;
    constexpr int CHUNKS = BN / 32 / 2;                              // 32-column chunks per warp
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int tm = tile / n_tn, tn = tile - tm * n_tn;
      const int t = wsi_tile_group(segs, tm);
      const int row0 = segs.ptr[t] + (tm - segs.tile_start[t]) * PAIR_M + (int)rank * BM + q * 32;
      const int rows_left = segs.ptr[t + 1] - (row0 + rsub);         // this lane handles rows row0 + rsub + 4 it
      // ---- plain epilogue through TMA stores (SASS UTMASTG): thread = accumulator row straight out of TMEM, + bias
      // (+ GELU), one swizzled 16 B store per 4 columns into the warp's 4 KB staging buffer, then ONE bulk tensor store of
      // the 32 x 32 chunk - no smem read-back, no per-lane global stores.  Warps whose 32 rows straddle the end of the
      // node type (the next rows belong to another tile) and the FULL epilogue take the generic path below.
      if (!FULL && a.tma_store && (TERMS == 1 || !a.y_op) && segs.ptr[t + 1] - row0 >= 32) {
        const int col_w = tn * BN + half * (BN / 2);                 // first column of this warp
        const float* bias_w = ep.bias ? ep.bias + (int64_t)t * ep.n_out : nullptr;
        const bool two = EPI_BUFS == 2 && !a.y_op;                   // y double-buffered (else buffer 1 stages y_op)
        mbar_wait(tfull_bar + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
        const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int col0 = col_w + c * 32;
          if (col0 >= a.n_out || (a.dbg & 4)) break;                 // warp-uniform
          float v[32];
          if (a.dbg & 16) {                                          // development: no TMEM read
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          } else
          tc_ld_32x32(taddr + c * 32, v);
          float4 bb[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            bb[j] = (bias_w && col0 + 4 * j < a.n_out) ? __ldg(reinterpret_cast<const float4*>(bias_w + col0 + 4 * j)) : zero4;
          tc_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j] += bb[j].x; v[4 * j + 1] += bb[j].y; v[4 * j + 2] += bb[j].z; v[4 * j + 3] += bb[j].w;
            if (GELU) { v[4 * j] = wsi_gelu(v[4 * j]); v[4 * j + 1] = wsi_gelu(v[4 * j + 1]); v[4 * j + 2] = wsi_gelu(v[4 * j + 2]); v[4 * j + 3] = wsi_gelu(v[4 * j + 3]); }
          }
          // the staging buffer is free once the bulk store that last read it has finished READING smem
          if (lane == 0) { if (two) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
          __syncwarp();
          if (ep.y) {
            const uint32_t sb = stg_u32 + (two ? store_buf * (EPI_WARP_FLOATS * 4) : 0) + (uint32_t)lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((((uint32_t)j) ^ sw) << 4)), "f"(v[4 * j]),
                           "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
          }
          if (TERMS == 1 && a.y_op) {                                // 16-bit copy: two chunks fill one 128 B staging row
            const uint32_t sb = stg_u32 + EPI_WARP_FLOATS * 4 + (uint32_t)lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint2 lo = a.op_bf16 ? pack4_bf16(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]))
                                   : pack4_f16(make_float4(v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3]));
              uint2 hi = a.op_bf16 ? pack4_bf16(make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]))
                                   : pack4_f16(make_float4(v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + ((((uint32_t)((c & 1) * 4 + j)) ^ sw) << 4)),
                           "r"(lo.x), "r"(lo.y), "r"(hi.x), "r"(hi.y) : "memory");
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && !(a.dbg & 8)) {
            if (ep.y) tma_store_2d(&tmY, stg_u32 + (two ? store_buf * (EPI_WARP_FLOATS * 4) : 0), col0, row0);
            if (TERMS == 1 && a.y_op && ((c & 1) || col0 + 32 >= a.n_out))
              tma_store_2d(&tmY16, stg_u32 + EPI_WARP_FLOATS * 4, col0 & ~63, row0);
            bulk_commit();
          }
          store_buf ^= 1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa(tempty_bar + 8 * acc, 0));   // on the leader's barrier
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if (a.tma_store) {                                             // the generic path reuses the staging buffer: drain first
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
      }
      const int n0 = tn * BN + half * (BN / 2) + c4;                 // this lane's first column
      const float alpha = (FULL && ep.skip) ? wsi_sigmoid(__ldg(ep.skip + t)) : 1.0f;
      float* yp = ep.y ? ep.y + (int64_t)(row0 + rsub) * ep.ldy + n0 : nullptr;
      const float* bias_p = ep.bias ? ep.bias + (int64_t)t * ep.n_out + n0 : nullptr;
      const float* res_p = (FULL && ep.skip) ? ep.res + (int64_t)(row0 + rsub) * ep.ldres + n0 : nullptr;
      const float* mask_p = (FULL && ep.drop_mask) ? ep.drop_mask + (int64_t)(row0 + rsub) * ep.ldmask + n0 : nullptr;
      float gate[8], rscl[8];
      if (FULL) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const bool ok = it * 4 < rows_left;
          gate[it] = (ep.row_gate && ok) ? __ldg(ep.row_gate + row0 + rsub + it * 4) : 1.f;
          rscl[it] = (ep.row_scale && ok) ? __ldg(ep.row_scale + row0 + rsub + it * 4) : 1.f;
        }
      }
      // residual rows of chunk 0 are requested BEFORE waiting for the accumulator, those of chunk c+1 while chunk c is
      // processed: with one tile per CTA pair (the a_linear shapes) the epilogue is exposed and was bound by the
      // latency of these loads, 4 KB per warp in flight
      float4 rr[2][8];
      auto load_res = [&](int c, float4* dst) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const bool ok = n0 + c * 32 < a.n_out && it * 4 < rows_left;
          dst[it] = (res_p && ok) ? __ldg(reinterpret_cast<const float4*>(res_p + (int64_t)it * 4 * ep.ldres + c * 32)) : zero4;
        }
      };
      if (FULL) load_res(0, rr[0]);
      // the bias vectors of all chunks too: a load issued inside the chunk loop sat in the dependent chain of every chunk
      float4 bb_all[CHUNKS];
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c)
        bb_all[c] = (bias_p && n0 + c * 32 < a.n_out) ? __ldg(reinterpret_cast<const float4*>(bias_p + c * 32)) : zero4;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * (BN / 2));
#pragma unroll
      for (int c = 0; c < CHUNKS; ++c) {
        if (n0 - c4 + c * 32 >= a.n_out || rows_left + rsub <= 0 || (a.dbg & 4)) break;   // warp-uniform: nothing left to store
        float v[32];
        if (a.dbg & 16) {                                            // development: no TMEM read
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        } else {
          tc_ld_32x32(taddr + c * 32, v);
        }
        const bool n_ok = n0 + c * 32 < a.n_out;
        // issue every global load of this chunk before waiting on TMEM
        const float4 bb = bb_all[c];
        if (FULL && c + 1 < CHUNKS) load_res(c + 1, rr[(c + 1) & 1]);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(stg + lane * EPI_LD + ((j ^ (lane & 7)) << 2)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          if (n_ok && it * 4 < rows_left ) {
            const float4 accv = *reinterpret_cast<const float4*>(stg + (it * 4 + rsub) * EPI_LD + ((((lane & 7) ^ ((it * 4 + rsub) & 7))) << 2));
            // (the dropout mask - training only - is read where it is used: prefetching it next to the residual cost 32
            //  registers and pushed the FULL epilogue into local-memory spills)
            const float4 mk = (FULL && mask_p) ? __ldg(reinterpret_cast<const float4*>(mask_p + (int64_t)it * 4 * ep.ldmask + c * 32)) : one4;
            const float4 o = FULL ? epi_mix4<true, GELU>(ep, accv, bb, mk, rr[c & 1][it], alpha, gate[it] != 0.f, rscl[it])
                                  : epi_mix4<false, GELU>(ep, accv, bb, one4, zero4, 1.f, true, 1.f);
            if (ep.y && !(a.dbg & 8)) *reinterpret_cast<float4*>(yp + (int64_t)it * 4 * ep.ldy + c * 32) = o;
            else if (a.dbg & 8) { if (o.x == 123456.789f) yp[0] = o.y; }       // development: no global stores (keep the math alive)
            if (a.y_op && !(a.dbg & 8)) {
              uint16_t* ys = reinterpret_cast<uint16_t*>(a.y_op) + (int64_t)(row0 + rsub + it * 4) * a.n_out + n0 + c * 32;
              if (TERMS == 3) store_operand4<WSI_OPF_BF16X3>(ys, (int64_t)a.n_rows * a.n_out, o);
              else if (a.op_bf16) store_operand4<WSI_OPF_BF16>(ys, 0, o);
              else store_operand4<WSI_OPF_F16>(ys, 0, o);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa(tempty_bar + 8 * acc, 0));   // on the leader's barrier
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait_all();                       // this thread's TMA stores have landed
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                                    // nobody touches the peer's smem / TMEM after this
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------ host side
inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int terms_of(int opf) { return opf == WSI_OPF_BF16X3 ? 2 : 1; }   // 16-bit matrices per operand

}  // namespace

// Shapes the tensor-core path takes: K a multiple of 8 (16 B rows for TMA / float4 conversion), output rows 16 B
// aligned, and enough work to fill 128-row MMA tiles (small [B, D] readout GEMMs stay on the SIMT path).
bool wsi_typed_linear_tc_supported(int64_t n_rows, int K, int n_out, int64_t ldx) {
  // (n_out % 4 != 0 is accepted here; the callers then require a padded output pitch and no per-column vectors:
  //  the epilogue stores whole float4 groups, see wsi_typed_linear_f32)
  return n_rows >= 512 && n_rows < (1ll << 30) && K >= 64 && K % 8 == 0 && n_out >= 64 &&
         ldx % 4 == 0 && (int64_t)n_out * WSI_MAX_TYPES < (1ll << 30);
}

bool wsi_opf_valid(int opf) { return opf == WSI_OPF_BF16X3 || opf == WSI_OPF_F16 || opf == WSI_OPF_BF16; }

int64_t wsi_typed_linear_tc_workspace(int64_t n_rows, int K, int n_out, int T, int opf) {
  const int64_t m = terms_of(opf);
  return align256(m * n_rows * K * 2) + align256(m * (int64_t)T * n_out * K * 2) + 1024;
}

// GEMM on operands already in operand form (`opf`): a_ws / w_ws 16-bit [m * n_rows, K] / [m * T * n_out, K], m = 2 for
// the [hi; lo] split (hi rows, then lo rows), else 1.
int wsi_typed_linear_tc_gemm(const void* a_ws, const void* w_ws, int K, const int32_t* type_ptr_host, int T,
                             const LinearEpilogue& ep, void* y_op, int opf, cudaStream_t stream) {
  const int64_t n_rows = type_ptr_host[T];
  const int n_out = ep.n_out;
  WSI_CHECK_ARG(wsi_opf_valid(opf), "typed_linear(tcgen05): unknown operand format %d", opf);
  WSI_CHECK_ARG((ep.y || y_op) && (!ep.y || (aligned16(ep.y) && ep.ldy % 4 == 0)),
                "typed_linear(tcgen05): y must be 16 B aligned with a row stride that is a multiple of 4 floats");
  WSI_CHECK_ARG((reinterpret_cast<uintptr_t>(a_ws) & 127) == 0 && (reinterpret_cast<uintptr_t>(w_ws) & 127) == 0 &&
                    (!y_op || (reinterpret_cast<uintptr_t>(y_op) & 15) == 0),
                "typed_linear(tcgen05): operands must be 128 B aligned");
  WSI_CHECK_ARG((!ep.bias || aligned16(ep.bias)) && (!ep.drop_mask || (aligned16(ep.drop_mask) && ep.ldmask % 4 == 0)) &&
                    (!ep.res || (aligned16(ep.res) && ep.ldres % 4 == 0)),
                "typed_linear(tcgen05): bias / drop_mask / res must be 16 B aligned with row strides multiple of 4 floats");
  TypeSegs segs;
  if (wsi_make_segs(&segs, type_ptr_host, T, PAIR_M) != 0) { wsi_set_error("typed_linear: bad type_ptr"); return WSI_ERR_ARG; }
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    const void* fns[8] = {(const void*)typed_linear_tc_kernel<false, false, 3>, (const void*)typed_linear_tc_kernel<true, false, 3>,
                          (const void*)typed_linear_tc_kernel<false, true, 3>, (const void*)typed_linear_tc_kernel<true, true, 3>,
                          (const void*)typed_linear_tc_kernel<false, false, 1>, (const void*)typed_linear_tc_kernel<true, false, 1>,
                          (const void*)typed_linear_tc_kernel<false, true, 1>, (const void*)typed_linear_tc_kernel<true, true, 1>};
    for (int i = 0; i < 8 && attr_err == cudaSuccess; ++i)
      attr_err = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_of(i < 4 ? 3 : 1));
  });
  WSI_CHECK_CUDA(attr_err);

  const bool bf16 = opf != WSI_OPF_F16;
  const int m = terms_of(opf);
  const CUtensorMapDataType dt16 = bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tmA, tmB, tmY, tmY16;
  int rc = make_map(&tmA, a_ws, m * n_rows, K, (int64_t)K * 2, BM, BK, dt16);
  if (rc != WSI_OK) return rc;
  rc = make_map(&tmB, w_ws, m * (int64_t)T * n_out, K, (int64_t)K * 2, BN / 2, BK, dt16);
  if (rc != WSI_OK) return rc;
  const bool full = ep.skip || ep.drop_mask || ep.row_scale;
  // output maps of the TMA-store epilogue (plain epilogue only; column counts that are not a multiple of 4 - the padded
  // k-NN dot-product matrix - keep the generic stores, whose last float4 may spill into the row padding)
  const bool tma_store = !full && n_out % 4 == 0 && !wsi_dev()->tc_no_tma_store && (m == 1 || !y_op) && (!y_op || n_out % 8 == 0);
  tmY = tmA; tmY16 = tmA;
  if (tma_store) {
    if (ep.y) {
      rc = make_map(&tmY, ep.y, n_rows, n_out, ep.ldy * 4, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
      if (rc != WSI_OK) return rc;
    }
    if (y_op) {
      rc = make_map(&tmY16, y_op, n_rows, n_out, (int64_t)n_out * 2, 32, 64, dt16);
      if (rc != WSI_OK) return rc;
    }
  }
  TcArgs a{};
  a.tma_store = tma_store ? 1 : 0;
  a.n_rows = (int)n_rows; a.w_rows = T * n_out; a.K = K; a.n_out = n_out;
  a.n_tiles_m = segs.tile_start[T]; a.n_tiles_n = (n_out + BN - 1) / BN;
  a.y_op = y_op;
  a.idesc = idesc_of(bf16);
  a.op_bf16 = bf16 ? 1 : 0;
  a.dbg = wsi_dev()->tc_debug;
  const int total = a.n_tiles_m * a.n_tiles_n;
  if (total == 0) return WSI_OK;
  const int max_clusters = sms / 2;                                  // persistent: one CTA pair per two SMs
  const int clusters = total < max_clusters ? total : max_clusters;
  const bool gelu = ep.act == WSI_ACT_GELU;       // compile-time in the kernel: the erf code must not sit (predicated off) in the plain epilogue
  const dim3 grid(2 * clusters), block(THREADS);
  cudaError_t le;
#define TC_LAUNCH(F, G, TR) le = wsi_launch_pdl(typed_linear_tc_kernel<F, G, TR>, grid, block, smem_bytes_of(TR), stream, tmA, tmB, tmY, tmY16, segs, ep, a)
  if (m == 2) {
    if (full && gelu) TC_LAUNCH(true, true, 3); else if (full) TC_LAUNCH(true, false, 3);
    else if (gelu) TC_LAUNCH(false, true, 3); else TC_LAUNCH(false, false, 3);
  } else {
    if (full && gelu) TC_LAUNCH(true, true, 1); else if (full) TC_LAUNCH(true, false, 1);
    else if (gelu) TC_LAUNCH(false, true, 1); else TC_LAUNCH(false, false, 1);
  }
#undef TC_LAUNCH
  WSI_CHECK_CUDA(le);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

int wsi_convert_colsum_launch(const float* src, int64_t ld, int K, const int32_t* type_ptr_host, int T, void* dst,
                              float* colsum, float* partial, cudaStream_t stream) {
  TypeSegs segs;
  if (wsi_make_segs(&segs, type_ptr_host, T, CS_ROWS) != 0) { wsi_set_error("to_operand_colsum: bad type_ptr"); return WSI_ERR_ARG; }
  const int tiles = segs.tile_start[T];
  const int64_t n_rows = type_ptr_host[T];
  if (tiles > 0) {
    convert_colsum_kernel<<<tiles, 256, 0, stream>>>(src, ld, n_rows, K, segs, reinterpret_cast<uint16_t*>(dst), partial);
    WSI_CHECK_LAUNCH();
  }
  colsum_finish_kernel<<<dim3((K + 63) / 64, T), 256, 0, stream>>>(partial, K, segs, colsum);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

int64_t wsi_convert_colsum_tiles(const int32_t* type_ptr_host, int T) {
  int64_t tiles = 0;
  for (int t = 0; t < T; ++t) tiles += (type_ptr_host[t + 1] - type_ptr_host[t] + CS_ROWS - 1) / CS_ROWS;
  return tiles;
}

int wsi_gather_rows16_launch(const void* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, void* dst,
                             cudaStream_t stream) {
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  const int64_t total = rows * (K / 8);
  int blocks = (int)((total + 255) / 256);
  if (blocks > sms * 16) blocks = sms * 16;
  gather_rows16_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(src), ld_src / 8, row_idx, rows, K / 8,
                                                   reinterpret_cast<uint4*>(dst));
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

// fp32 -> operand form of up to two row-strided matrices in one launch (b.rows == 0: only a)
int wsi_split_launch(const float* a_src, int64_t a_ld, int64_t a_rows, void* a_dst, const float* b_src, int64_t b_ld,
                     int64_t b_rows, void* b_dst, int K, int opf, cudaStream_t stream, const int32_t* a_row_idx) {
  WSI_CHECK_ARG(wsi_opf_valid(opf), "operand conversion: unknown operand format %d", opf);
  WSI_CHECK_ARG(K % 8 == 0 && a_ld % 4 == 0 && (b_rows == 0 || b_ld % 4 == 0) && aligned16(a_src) && aligned16(b_src) &&
                    (reinterpret_cast<uintptr_t>(a_dst) & 7) == 0 && (reinterpret_cast<uintptr_t>(b_dst) & 7) == 0,
                "operand conversion: K must be a multiple of 8, rows 16 B aligned");
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  SplitJob ja{a_src, a_ld, a_rows, a_dst, a_row_idx};
  SplitJob jb{b_src, b_ld, b_rows, b_dst, nullptr};
  const int64_t groups = (ja.rows + jb.rows) * (K / 4);
  if (groups == 0) return WSI_OK;
  int sblocks = (int)((groups + 1023) / 1024);
  if (sblocks > sms * 8) sblocks = sms * 8;
  if (sblocks < 1) sblocks = 1;
  if (opf == WSI_OPF_BF16X3) convert_operand_kernel<WSI_OPF_BF16X3><<<sblocks, 256, 0, stream>>>(ja, jb, K);
  else if (opf == WSI_OPF_F16) convert_operand_kernel<WSI_OPF_F16><<<sblocks, 256, 0, stream>>>(ja, jb, K);
  else convert_operand_kernel<WSI_OPF_BF16><<<sblocks, 256, 0, stream>>>(ja, jb, K);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

int wsi_typed_linear_tc_launch(const float* x, int64_t ldx, const float* w, int K, const int32_t* type_ptr_host,
                               int T, const LinearEpilogue& ep, int opf, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream) {
  const int64_t n_rows = type_ptr_host[T];
  const int n_out = ep.n_out;
  WSI_CHECK_ARG(workspace && workspace_bytes >= wsi_typed_linear_tc_workspace(n_rows, K, n_out, T, opf),
                "typed_linear(tcgen05): workspace of %lld bytes needed", (long long)wsi_typed_linear_tc_workspace(n_rows, K, n_out, T, opf));
  uintptr_t wsp = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023;
  void* a_ws = reinterpret_cast<void*>(wsp);
  void* w_ws = reinterpret_cast<void*>(wsp + align256(terms_of(opf) * n_rows * K * 2));
  int rc = wsi_split_launch(x, ldx, n_rows, a_ws, w, K, (int64_t)T * n_out, w_ws, K, opf, stream, nullptr);
  if (rc != WSI_OK) return rc;
  return wsi_typed_linear_tc_gemm(a_ws, w_ws, K, type_ptr_host, T, ep, nullptr, opf, stream);
}
