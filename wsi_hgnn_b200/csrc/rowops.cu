// Row-wise HBM-bound kernels: typed readout (K4), typed LayerNorm, HGT segment combine.
#include "operand.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// Typed readout: dgl.readout.{sum,mean,max}_nodes(graph, 'h', ntype=) (reference pooling/avg_pooling.py:15-17,
// sum_pooling.py:14-16, max_pooling.py:15-17).  grid = (segment, 128-column chunk, row split); every CTA
// reduces a contiguous slab of rows with coalesced float4 loads; partials of split segments go through the
// caller's workspace and are combined by a second tiny kernel (deterministic, no atomics).
constexpr int POOL_THREADS = 256;   // 8 warps; a warp covers 128 columns, warps stride over rows

__device__ __forceinline__ float4 pool_comb(float4 a, float4 b, int op) {
  if (op == WSI_POOL_MAX) return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__global__ void __launch_bounds__(POOL_THREADS)
segment_pool_kernel(const float* __restrict__ x, int64_t ldx, const int* __restrict__ seg_ptr, int D, int op,
                    int n_split, float* __restrict__ out, int64_t ldo, float* __restrict__ partial, int force_partial,
                    const float* __restrict__ aff_M, int aff_n_out, int aff_B, float* __restrict__ aff_pdots) {
  __shared__ float4 sm[POOL_THREADS / 32][32];
  const int seg = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = (blockIdx.y * 32 + lane) * 4;
  const int beg = seg_ptr[seg], end = seg_ptr[seg + 1];
  const int n = end - beg;
  const int per = (n + n_split - 1) / n_split;
  const int r0 = beg + blockIdx.z * per;
  const int r1 = min(end, r0 + per);
  const float ident = op == WSI_POOL_MAX ? -INFINITY : 0.f;
  float4 acc = make_float4(ident, ident, ident, ident);
  const bool vec = (col + 3 < D) && (ldx % 4 == 0);
  constexpr int W = POOL_THREADS / 32;
  int r = r0 + warp;
  if (vec) {                                            // 4 rows in flight per thread
    for (; r + 3 * W < r1; r += 4 * W) {
      const float* p = x + (int64_t)r * ldx + col;
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(p));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * ldx));
      const float4 v2 = __ldg(reinterpret_cast<const float4*>(p + (int64_t)2 * W * ldx));
      const float4 v3 = __ldg(reinterpret_cast<const float4*>(p + (int64_t)3 * W * ldx));
      acc = pool_comb(pool_comb(acc, v0, op), pool_comb(pool_comb(v1, v2, op), v3, op), op);
    }
  }
  for (; r < r1; r += W) {
    const float* p = x + (int64_t)r * ldx + col;
    float4 v;
    if (vec) v = __ldg(reinterpret_cast<const float4*>(p));
    else {
      v.x = col < D ? __ldg(p) : ident;         v.y = col + 1 < D ? __ldg(p + 1) : ident;
      v.z = col + 2 < D ? __ldg(p + 2) : ident; v.w = col + 3 < D ? __ldg(p + 3) : ident;
    }
    acc = pool_comb(acc, v, op);
  }
  sm[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < POOL_THREADS / 32; ++w) acc = pool_comb(acc, sm[w][lane], op);
    float vals[4] = {acc.x, acc.y, acc.z, acc.w};
    if (aff_pdots) {
      // sum / mean readout followed by an affine map: the map is linear in the slab sums, so this CTA contributes
      // its own <M[t, o, cols], slab sum> and the [.., D] partial never goes to memory
      const int t = seg / aff_B;
      float* pd = aff_pdots + (((int64_t)blockIdx.z * gridDim.x + seg) * gridDim.y + blockIdx.y) * aff_n_out;
      for (int o = 0; o < aff_n_out; ++o) {
        const float* mrow = aff_M + ((int64_t)t * aff_n_out + o) * D + col;
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (col + c < D) d = fmaf(vals[c], __ldg(mrow + c), d);
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sft);
        if (lane == 0) pd[o] = d;
      }
      return;
    }
    if (n_split == 1 && !force_partial) {
      const float scale = op == WSI_POOL_MEAN ? 1.f / (float)max(n, 1) : 1.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (col + c < D) out[(int64_t)seg * ldo + col + c] = n > 0 ? vals[c] * scale : 0.f;
    } else {
      float* p = partial + ((int64_t)blockIdx.z * gridDim.x + seg) * D;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (col + c < D) p[col + c] = vals[c];
    }
  }
}

__global__ void segment_pool_finish_kernel(const float* __restrict__ partial, const int* __restrict__ seg_ptr,
                                           int64_t n_seg, int D, int op, int n_split, float* __restrict__ out,
                                           int64_t ldo) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_seg * D) return;
  int64_t seg = idx / D;
  int c = (int)(idx % D);
  int n = seg_ptr[seg + 1] - seg_ptr[seg];
  float acc = op == WSI_POOL_MAX ? -INFINITY : 0.f;
#pragma unroll 8
  for (int z = 0; z < n_split; ++z) {
    float v = partial[((int64_t)z * n_seg + seg) * D + c];
    acc = op == WSI_POOL_MAX ? fmaxf(acc, v) : acc + v;
  }
  if (op == WSI_POOL_MEAN) acc /= (float)max(n, 1);
  out[seg * ldo + c] = n > 0 ? acc : 0.f;
}

// Fused finish of the typed readout when what follows the pooling is affine and narrow (HEATNet2's
// sum_t linears_prediction[t](pool_t), models/HEATNet2.py:181-194; HGT's per-layer readout, models/HGT.py:189-199; and
// HEATNet4's linears_prediction -> cat -> head_2 -> head_1 -> head, which has no nonlinearity
// (models/HEATNet4.py:216-245) and is collapsed on the host into one [out, D] map per node type):
//   out[b, o] (+)= b_total[o] + sum_t scale[t*B + b] * ( <M[t, o, :], pool(t, b)> + c[t, o] )
// one block per graph b: threads own columns, reduce the pooling partials of their column, multiply by M, block-reduce.
constexpr int AFF_MAX_OUT = 8;

// sum / mean: the slab dots of segment_pool_kernel -> out.  pdots [n_split, T*B, chunks, n_out]
__global__ void __launch_bounds__(128)
segment_pool_affine_dots_finish_kernel(const float* __restrict__ pdots, const int* __restrict__ seg_ptr, int T, int B,
                                       int chunks, int op, int n_split, const float* __restrict__ c,
                                       const float* __restrict__ b_total, const float* __restrict__ seg_scale, int n_out,
                                       int accumulate, float* __restrict__ out, int64_t ldo) {
  __shared__ float red[4][AFF_MAX_OUT];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_seg = (int64_t)T * B;
  float tot[AFF_MAX_OUT];
#pragma unroll
  for (int o = 0; o < AFF_MAX_OUT; ++o) tot[o] = 0.f;
  for (int t = 0; t < T; ++t) {
    const int64_t seg = (int64_t)t * B + b;
    const int n = seg_ptr[seg + 1] - seg_ptr[seg];
    float sc = seg_scale ? seg_scale[seg] : 1.f;
    if (n == 0 || sc == 0.f) continue;                  // empty type: zero block (models/HEATNet4.py:240)
    const float pool_scale = op == WSI_POOL_MEAN ? 1.f / (float)n : 1.f;
    for (int i = threadIdx.x; i < n_split * chunks; i += blockDim.x) {
      const int z = i / chunks, ch = i - z * chunks;
      const float* pd = pdots + (((int64_t)z * n_seg + seg) * chunks + ch) * n_out;
#pragma unroll
      for (int o = 0; o < AFF_MAX_OUT; ++o)
        if (o < n_out) tot[o] = fmaf(sc * pool_scale, pd[o], tot[o]);
    }
    if (threadIdx.x == 0 && c)
#pragma unroll
      for (int o = 0; o < AFF_MAX_OUT; ++o)
        if (o < n_out) tot[o] = fmaf(sc, __ldg(c + (int64_t)t * n_out + o), tot[o]);
  }
#pragma unroll
  for (int o = 0; o < AFF_MAX_OUT; ++o) {
    float v = tot[o];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][o] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_out) {
    float v = b_total ? b_total[threadIdx.x] : 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    float* p = out + (int64_t)b * ldo + threadIdx.x;
    *p = accumulate ? *p + v : v;
  }
}

// max: not linear - reduce the [.., D] partials first
__global__ void __launch_bounds__(256)
segment_pool_affine_finish_kernel(const float* __restrict__ partial, const int* __restrict__ seg_ptr, int T, int B, int D,
                                  int op, int n_split, const float* __restrict__ M, const float* __restrict__ c,
                                  const float* __restrict__ b_total, const float* __restrict__ seg_scale, int n_out,
                                  int accumulate, float* __restrict__ out, int64_t ldo) {
  __shared__ float red[8][AFF_MAX_OUT];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_seg = (int64_t)T * B;
  float tot[AFF_MAX_OUT];
#pragma unroll
  for (int o = 0; o < AFF_MAX_OUT; ++o) tot[o] = 0.f;
  for (int t = 0; t < T; ++t) {
    const int64_t seg = (int64_t)t * B + b;
    const int n = seg_ptr[seg + 1] - seg_ptr[seg];
    const float sc = seg_scale ? seg_scale[seg] : 1.f;
    if (n == 0 || sc == 0.f) continue;                  // empty type: zero block (models/HEATNet4.py:240)
    float dot[AFF_MAX_OUT];
#pragma unroll
    for (int o = 0; o < AFF_MAX_OUT; ++o) dot[o] = 0.f;
    for (int col = threadIdx.x; col < D; col += blockDim.x) {
      float acc = op == WSI_POOL_MAX ? -INFINITY : 0.f;
#pragma unroll 8
      for (int z = 0; z < n_split; ++z) {
        const float v = partial[((int64_t)z * n_seg + seg) * D + col];
        acc = op == WSI_POOL_MAX ? fmaxf(acc, v) : acc + v;
      }
      if (op == WSI_POOL_MEAN) acc /= (float)n;
#pragma unroll
      for (int o = 0; o < AFF_MAX_OUT; ++o)
        if (o < n_out) dot[o] = fmaf(__ldg(M + ((int64_t)t * n_out + o) * D + col), acc, dot[o]);
    }
#pragma unroll
    for (int o = 0; o < AFF_MAX_OUT; ++o) tot[o] = fmaf(sc, dot[o], tot[o]);
    if (threadIdx.x == 0 && c)
#pragma unroll
      for (int o = 0; o < AFF_MAX_OUT; ++o)
        if (o < n_out) tot[o] = fmaf(sc, __ldg(c + (int64_t)t * n_out + o), tot[o]);
  }
#pragma unroll
  for (int o = 0; o < AFF_MAX_OUT; ++o) {
    float v = tot[o];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[warp][o] = v;
  }
  __syncthreads();
  if (threadIdx.x < n_out) {
    float v = b_total ? b_total[threadIdx.x] : 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    float* p = out + (int64_t)b * ldo + threadIdx.x;
    *p = accumulate ? *p + v : v;
  }
}

int pool_splits(int64_t n_rows, int64_t n_seg, int D) {
  // enough CTAs to fill the chip even when one (type, graph) segment holds all the rows
  int chunks = (D + 127) / 128;
  int64_t ctas = n_seg * chunks;
  const int sms = wsi_num_sms();
  int64_t want = (int64_t)(sms > 0 ? sms : 148) * 4;
  int s = (int)((want + ctas - 1) / ctas);
  int64_t max_by_rows = (n_rows + 255) / 256;     // at least ~256 rows per slab
  if (s > max_by_rows) s = (int)max_by_rows;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

// ------------------------------------------------------------------------------------------------
// Typed LayerNorm (reference models/HGT.py:123-124): one warp per row, two-pass mean / variance.
__global__ void __launch_bounds__(256)
typed_layernorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ row_gate, TypeSegs segs, int D, float eps,
                       float* __restrict__ y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int n_rows = segs.ptr[segs.T];
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 8) {
    int t = 0;
    while (t + 1 < segs.T && row >= segs.ptr[t + 1]) ++t;
    const float* xr = x + (int64_t)row * ldx;
    if (row_gate && __ldg(row_gate + row) == 0.f) {       // passthrough row (no incoming relation): not normalised
      if (y != x || ldy != ldx) {
        float* yr = y + (int64_t)row * ldy;
        for (int c = lane; c < D; c += 32) yr[c] = xr[c];
      }
      continue;
    }
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += xr[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    const float rstd = rsqrtf(v / (float)D + eps);
    const float* g = gamma + (int64_t)t * D;
    const float* b = beta + (int64_t)t * D;
    float* yr = y + (int64_t)row * ldy;
    for (int c = lane; c < D; c += 32) yr[c] = (xr[c] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
  }
}

// Same, D = 128 * NV: the row lives in registers (one 16-byte load per lane and 128 columns, read ONCE), and the
// operand-form copy of the result - the A operand of the next layer's K | V | Q GEMMs - is written in the same pass.
template <int NV>
__global__ void __launch_bounds__(256)
typed_layernorm_vec_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                           const float* __restrict__ beta, const float* __restrict__ row_gate, TypeSegs segs, float eps,
                           float* __restrict__ y, int64_t ldy, uint16_t* __restrict__ y_op, int64_t lo_off) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int n_rows = segs.ptr[segs.T];
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 8) {
    int t = 0;
    while (t + 1 < segs.T && row >= segs.ptr[t + 1]) ++t;
    const float* xr = x + (int64_t)row * ldx;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    if (!(row_gate && __ldg(row_gate + row) == 0.f)) {     // (gate == 0: passthrough row, copied through un-normalised)
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      const float mean = s / (float)D;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q); q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(FULL, q, o);
      const float rstd = rsqrtf(q / (float)D + eps);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + (int64_t)t * D + (i * 32 + lane) * 4));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + (int64_t)t * D + (i * 32 + lane) * 4));
        v[i].x = fmaf(v[i].x * rstd, g.x, b.x); v[i].y = fmaf(v[i].y * rstd, g.y, b.y);
        v[i].z = fmaf(v[i].z * rstd, g.z, b.z); v[i].w = fmaf(v[i].w * rstd, g.w, b.w);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (y) *reinterpret_cast<float4*>(y + (int64_t)row * ldy + (i * 32 + lane) * 4) = v[i];
      if (y_op) store_operand4_rt(y_op + (int64_t)row * D + (i * 32 + lane) * 4, lo_off, v[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// HGT: agg[v] = inv_r[v] * sum of the messages of the (v, relation) segments of row v
// (stack->mean of multi_update_all(..., cross_reducer='mean'), reference models/HGT.py:105-106).
// seg_pos (optional): msg row of segment s is seg_pos[s] (the relation-sorted order of the tensor-core transforms);
// agg_op (optional): the operand-form copy for the a_linear GEMM (lo_off as in store_operand4_rt).
__device__ __forceinline__ float4 ld_msg4(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld_msg4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// 16-bit message storage (fp16 / bf16 rows straight out of the relation_msg GEMM), 4 columns per lane
template <typename MT>
__global__ void __launch_bounds__(256)
segment_combine16_kernel(const MT* __restrict__ msg, int64_t ldm, const int* __restrict__ row_seg_ptr,
                         const int* __restrict__ seg_pos, const float* __restrict__ inv_r, int n_rows, int D,
                         float* __restrict__ agg, int64_t ldo, uint16_t* __restrict__ agg_op, int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 8) {
    const int s0 = row_seg_ptr[row], s1 = row_seg_ptr[row + 1];
    const float ir = inv_r[row];
    for (int c = lane * 4; c < D; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = s0; s < s1; ++s) {
        const int64_t r = seg_pos ? (int64_t)__ldg(seg_pos + s) : s;
        const float4 m = ld_msg4(msg + r * ldm + c);
        acc.x += m.x; acc.y += m.y; acc.z += m.z; acc.w += m.w;
      }
      acc.x *= ir; acc.y *= ir; acc.z *= ir; acc.w *= ir;
      if (agg) *reinterpret_cast<float4*>(agg + (int64_t)row * ldo + c) = acc;
      if (agg_op) store_operand4_rt(agg_op + (int64_t)row * D + c, lo_off, acc);
    }
  }
}

template <bool VEC4>
__global__ void __launch_bounds__(256)
segment_combine_kernel(const float* __restrict__ msg, int64_t ldm, const int* __restrict__ row_seg_ptr,
                       const int* __restrict__ seg_pos, const float* __restrict__ inv_r, int n_rows, int D,
                       float* __restrict__ agg, int64_t ldo, uint16_t* __restrict__ agg_op, int64_t lo_off) {
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * 8) {
    const int s0 = row_seg_ptr[row], s1 = row_seg_ptr[row + 1];
    const float ir = inv_r[row];
    if (VEC4) {
      for (int c = lane * 4; c < D; c += 128) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = s0; s < s1; ++s) {
          const int64_t r = seg_pos ? (int64_t)__ldg(seg_pos + s) : s;
          const float4 m = __ldg(reinterpret_cast<const float4*>(msg + r * ldm + c));
          acc.x += m.x; acc.y += m.y; acc.z += m.z; acc.w += m.w;
        }
        acc.x *= ir; acc.y *= ir; acc.z *= ir; acc.w *= ir;
        if (agg) *reinterpret_cast<float4*>(agg + (int64_t)row * ldo + c) = acc;
        if (agg_op) store_operand4_rt(agg_op + (int64_t)row * D + c, lo_off, acc);
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        float acc = 0.f;
        for (int s = s0; s < s1; ++s) acc += __ldg(msg + (seg_pos ? (int64_t)__ldg(seg_pos + s) : (int64_t)s) * ldm + c);
        agg[(int64_t)row * ldo + c] = acc * ir;
      }
    }
  }
}

}  // namespace

extern "C" int64_t wsi_segment_pool_workspace_bytes(int64_t n_rows, int64_t n_seg, int D) {
  int s = pool_splits(n_rows, n_seg, D);
  return s > 1 ? (int64_t)s * n_seg * D * (int64_t)sizeof(float) : 0;
}

extern "C" int wsi_segment_pool_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t n_seg,
                                    int64_t n_rows, int D, int op, float* out, int64_t ldo, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  WSI_CHECK_ARG(op == WSI_POOL_SUM || op == WSI_POOL_MEAN || op == WSI_POOL_MAX, "segment_pool: unknown op %d", op);
  WSI_CHECK_ARG(n_seg >= 0 && n_seg < 65536ll * 32768 && D >= 1, "segment_pool: bad n_seg / D");
  if (n_seg == 0) return WSI_OK;
  WSI_CHECK_ARG(seg_ptr && out && (x || n_rows == 0), "segment_pool: null pointer");
  const int splits = pool_splits(n_rows, n_seg, D);
  WSI_CHECK_ARG(splits == 1 || (workspace && workspace_bytes >= wsi_segment_pool_workspace_bytes(n_rows, n_seg, D)),
                "segment_pool: workspace too small");
  cudaStream_t st = wsi_stream(stream);
  dim3 grid((unsigned)n_seg, (D + 127) / 128, splits);
  segment_pool_kernel<<<grid, POOL_THREADS, 0, st>>>(x, ldx, seg_ptr, D, op, splits, out, ldo, (float*)workspace, 0,
                                                     nullptr, 0, 1, nullptr);
  WSI_CHECK_LAUNCH();
  if (splits > 1) {
    int64_t total = n_seg * D;
    segment_pool_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const float*)workspace, seg_ptr,
                                                                                n_seg, D, op, splits, out, ldo);
    WSI_CHECK_LAUNCH();
  }
  return WSI_OK;
}

// workspace: max(1, splits) * T * B * D floats (wsi_segment_pool_affine_workspace_bytes)
extern "C" int64_t wsi_segment_pool_affine_workspace_bytes(int64_t n_rows, int64_t n_seg, int D) {
  int s = pool_splits(n_rows, n_seg, D);
  return (int64_t)s * n_seg * D * (int64_t)sizeof(float);
}

extern "C" int wsi_segment_pool_affine_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int T, int B,
                                           int64_t n_rows, int D, int op, const float* M, const float* c,
                                           const float* b_total, const float* seg_scale, int n_out, int accumulate,
                                           float* out, int64_t ldo, void* workspace, int64_t workspace_bytes,
                                           void* stream) {
  WSI_CHECK_ARG(op == WSI_POOL_SUM || op == WSI_POOL_MEAN || op == WSI_POOL_MAX, "segment_pool_affine: unknown op %d", op);
  WSI_CHECK_ARG(T >= 1 && B >= 0 && D >= 1 && (int64_t)T * B < 65536ll * 32768, "segment_pool_affine: bad T / B / D");
  WSI_CHECK_ARG(n_out >= 1 && n_out <= AFF_MAX_OUT, "segment_pool_affine: 1 <= n_out <= %d (got %d)", AFF_MAX_OUT, n_out);
  if (B == 0) return WSI_OK;
  WSI_CHECK_ARG(seg_ptr && M && out && (x || n_rows == 0), "segment_pool_affine: null pointer");
  const int64_t n_seg = (int64_t)T * B;
  const int splits = pool_splits(n_rows, n_seg, D);
  WSI_CHECK_ARG(workspace && workspace_bytes >= wsi_segment_pool_affine_workspace_bytes(n_rows, n_seg, D),
                "segment_pool_affine: workspace too small");
  cudaStream_t st = wsi_stream(stream);
  dim3 grid((unsigned)n_seg, (D + 127) / 128, splits);
  if (op != WSI_POOL_MAX) {                               // linear readout: per-slab dots, [.., D] partials never stored
    segment_pool_kernel<<<grid, POOL_THREADS, 0, st>>>(x, ldx, seg_ptr, D, op, splits, nullptr, 0, nullptr, 1, M, n_out, B,
                                                       (float*)workspace);
    WSI_CHECK_LAUNCH();
    segment_pool_affine_dots_finish_kernel<<<B, 128, 0, st>>>((const float*)workspace, seg_ptr, T, B, (int)grid.y, op, splits,
                                                             c, b_total, seg_scale, n_out, accumulate, out, ldo);
    WSI_CHECK_LAUNCH();
    return WSI_OK;
  }
  segment_pool_kernel<<<grid, POOL_THREADS, 0, st>>>(x, ldx, seg_ptr, D, op, splits, nullptr, 0, (float*)workspace, 1,
                                                     nullptr, 0, 1, nullptr);
  WSI_CHECK_LAUNCH();
  segment_pool_affine_finish_kernel<<<B, 256, 0, st>>>((const float*)workspace, seg_ptr, T, B, D, op, splits, M, c, b_total,
                                                      seg_scale, n_out, accumulate, out, ldo);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

extern "C" int wsi_typed_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta,
                                   const float* row_gate, const int32_t* type_ptr_host, int T, int D, float eps, float* y,
                                   int64_t ldy, void* y_op, int opf, void* stream) {
  WSI_CHECK_ARG(x && gamma && beta && (y || y_op) && type_ptr_host, "typed_layernorm: null pointer");
  TypeSegs segs;
  WSI_CHECK_ARG(wsi_make_segs(&segs, type_ptr_host, T, 1) == 0, "typed_layernorm: bad type_ptr (T=%d)", T);
  int n_rows = segs.ptr[T];
  if (n_rows == 0) return WSI_OK;
  int blocks = (n_rows + 7) / 8;
  { const int sms = wsi_num_sms(); if (sms <= 0) return WSI_ERR_CUDA; if (blocks > sms * 16) blocks = sms * 16; }
  const bool vec = D % 128 == 0 && D <= 1024 && ldx % 4 == 0 && (!y || ldy % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(beta) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_op) & 7) == 0;
  WSI_CHECK_ARG(!y_op || (vec && opf >= 0 && opf <= 2),
                "typed_layernorm: the operand-form output needs D %% 128 == 0, D <= 1024 and 16 B aligned rows");
  WSI_CHECK_ARG(vec || y, "typed_layernorm: null pointer");
  if (vec) {
    const int64_t lo_off = opf == WSI_OPF_BF16X3 ? (int64_t)n_rows * D : (opf == WSI_OPF_F16 ? 0 : -1);
    switch (D / 128) {
#define CASE(NV) case NV: typed_layernorm_vec_kernel<NV><<<blocks, 256, 0, wsi_stream(stream)>>>( \
        x, ldx, gamma, beta, row_gate, segs, eps, y, ldy, reinterpret_cast<uint16_t*>(y_op), lo_off); break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
  } else {
    typed_layernorm_kernel<<<blocks, 256, 0, wsi_stream(stream)>>>(x, ldx, gamma, beta, row_gate, segs, D, eps, y, ldy);
  }
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

extern "C" int wsi_segment_combine(const void* msg_, int msg_dtype, int64_t ldm, const int32_t* row_seg_ptr, const int32_t* seg_pos,
                                   const float* node_inv_r, int64_t n_rows, int D, float* agg, int64_t ldo, void* agg_op,
                                   int opf, void* stream) {
  WSI_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "segment_combine: bad n_rows");
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(row_seg_ptr && node_inv_r && (agg || agg_op), "segment_combine: null pointer");
  WSI_CHECK_ARG(msg_dtype >= 0 && msg_dtype <= 2, "segment_combine: unknown message storage type %d", msg_dtype);
  const float* msg = reinterpret_cast<const float*>(msg_);
  const bool vec4 = D % 4 == 0 && ldm % 4 == 0 && (!agg || ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(msg) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(agg) & 15) == 0 && (reinterpret_cast<uintptr_t>(agg_op) & 7) == 0;
  WSI_CHECK_ARG(!agg_op || (vec4 && opf >= 0 && opf <= 2), "segment_combine: the operand-form output needs D %% 4 == 0, 16 B aligned rows");
  WSI_CHECK_ARG(vec4 || agg, "segment_combine: null pointer");
  int blocks = (int)((n_rows + 7) / 8);
  { const int sms = wsi_num_sms(); if (sms <= 0) return WSI_ERR_CUDA; if (blocks > sms * 16) blocks = sms * 16; }
  const int64_t lo_off = opf == WSI_OPF_BF16X3 ? n_rows * (int64_t)D : (opf == WSI_OPF_F16 ? 0 : -1);
  if (msg_dtype != 0) {
    WSI_CHECK_ARG(D % 4 == 0 && ldm % 4 == 0 && (!agg || ldo % 4 == 0) && (reinterpret_cast<uintptr_t>(msg_) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(agg) & 15) == 0 && (reinterpret_cast<uintptr_t>(agg_op) & 7) == 0,
                  "segment_combine: 16-bit messages need D %% 4 == 0 and 8 B aligned rows");
    if (msg_dtype == 1)
      segment_combine16_kernel<__half><<<blocks, 256, 0, wsi_stream(stream)>>>(
          reinterpret_cast<const __half*>(msg_), ldm, row_seg_ptr, seg_pos, node_inv_r, (int)n_rows, D, agg, ldo,
          reinterpret_cast<uint16_t*>(agg_op), lo_off);
    else
      segment_combine16_kernel<__nv_bfloat16><<<blocks, 256, 0, wsi_stream(stream)>>>(
          reinterpret_cast<const __nv_bfloat16*>(msg_), ldm, row_seg_ptr, seg_pos, node_inv_r, (int)n_rows, D, agg, ldo,
          reinterpret_cast<uint16_t*>(agg_op), lo_off);
    WSI_CHECK_LAUNCH();
    return WSI_OK;
  }
  if (vec4)
    segment_combine_kernel<true><<<blocks, 256, 0, wsi_stream(stream)>>>(msg, ldm, row_seg_ptr, seg_pos, node_inv_r, (int)n_rows,
                                                                         D, agg, ldo, reinterpret_cast<uint16_t*>(agg_op), lo_off);
  else
    segment_combine_kernel<false><<<blocks, 256, 0, wsi_stream(stream)>>>(msg, ldm, row_seg_ptr, seg_pos, node_inv_r, (int)n_rows,
                                                                          D, agg, ldo, nullptr, 0);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward of the fused a_linear epilogue  out = drop(lin) * a + x * (1 - a),  a = sigmoid(skip[t]) on rows with an
// incoming relation, passthrough (out = x) on the others (models/HEATNet4.py:122-136), in ONE pass over the rows:
//   d_lin = dout * a * mask        (input of the a_linear dgrad / wgrad)
//   d_x   = dout * (1 - a)         (dout itself on passthrough rows)
//   d_alpha[t] += sum over live rows of type t of <dout, out - x> / a      (d out / d a = drop(lin) - x = (out - x) / a)
// `lin` is not needed (and never materialised by the forward).  One warp per row; d_alpha through one atomic per block.
namespace {
__global__ void __launch_bounds__(256)
skip_mix_bwd_kernel(const float* __restrict__ dout, int64_t ldd, const float* __restrict__ out, int64_t ldo,
                    const float* __restrict__ x, int64_t ldx, const float* __restrict__ mask, int64_t ldm,
                    const float* __restrict__ skip, const float* __restrict__ gate, TypeSegs segs, int D,
                    float* __restrict__ d_lin, int64_t ldl, float* __restrict__ d_x, int64_t lddx, float* __restrict__ d_alpha) {
  __shared__ float s_part[8];
  __shared__ int s_type[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_rows = segs.ptr[segs.T];
  for (int row0 = blockIdx.x * 8; row0 < n_rows; row0 += gridDim.x * 8) {
    const int row = row0 + warp;
    float part = 0.f;
    int t = -1;
    if (row < n_rows) {
      t = 0;
      while (t + 1 < segs.T && row >= segs.ptr[t + 1]) ++t;
      const bool live = gate == nullptr || __ldg(gate + row) != 0.f;
      const float a = live ? 1.f / (1.f + expf(-__ldg(skip + t))) : 0.f;
      const float* dr = dout + (int64_t)row * ldd;
      const float* orow = out + (int64_t)row * ldo;
      const float* xr = x + (int64_t)row * ldx;
      for (int c = lane * 4; c < D; c += 128) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dr + c));
        float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
        if (mask) m = __ldg(reinterpret_cast<const float4*>(mask + (int64_t)row * ldm + c));
        *reinterpret_cast<float4*>(d_lin + (int64_t)row * ldl + c) = make_float4(g.x * a * m.x, g.y * a * m.y, g.z * a * m.z, g.w * a * m.w);
        *reinterpret_cast<float4*>(d_x + (int64_t)row * lddx + c) = make_float4(g.x * (1.f - a), g.y * (1.f - a), g.z * (1.f - a), g.w * (1.f - a));
        if (live) {
          const float4 o = __ldg(reinterpret_cast<const float4*>(orow + c));
          const float4 xv = __ldg(reinterpret_cast<const float4*>(xr + c));
          part = fmaf(g.x, o.x - xv.x, fmaf(g.y, o.y - xv.y, fmaf(g.z, o.z - xv.z, fmaf(g.w, o.w - xv.w, part))));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
      part = live ? part / a : 0.f;
    }
    if (lane == 0) { s_part[warp] = part; s_type[warp] = t; }
    __syncthreads();
    if (threadIdx.x == 0) {                               // rows of a block are consecutive: at most a few types per block
      int cur = -1;
      float acc = 0.f;
      for (int w = 0; w < 8; ++w) {
        if (s_type[w] < 0) continue;
        if (s_type[w] != cur) {
          if (cur >= 0 && acc != 0.f) atomicAdd(d_alpha + cur, acc);
          cur = s_type[w]; acc = 0.f;
        }
        acc += s_part[w];
      }
      if (cur >= 0 && acc != 0.f) atomicAdd(d_alpha + cur, acc);
    }
    __syncthreads();
  }
}
}  // namespace

extern "C" int wsi_skip_mix_bwd(const float* dout, int64_t ldd, const float* out, int64_t ldo, const float* x, int64_t ldx,
                                const float* drop_mask, int64_t ldm, const float* skip, const float* row_gate,
                                const int32_t* type_ptr_host, int T, int D, float* d_lin, int64_t ldl, float* d_x, int64_t lddx,
                                float* d_alpha, void* stream) {
  WSI_CHECK_ARG(dout && out && x && skip && type_ptr_host && d_lin && d_x && d_alpha, "skip_mix_bwd: null pointer");
  TypeSegs segs;
  WSI_CHECK_ARG(wsi_make_segs(&segs, type_ptr_host, T, 1) == 0, "skip_mix_bwd: bad type_ptr (T=%d)", T);
  WSI_CHECK_ARG(D >= 4 && D % 4 == 0 && ldd % 4 == 0 && ldo % 4 == 0 && ldx % 4 == 0 && ldl % 4 == 0 && lddx % 4 == 0 &&
                    (!drop_mask || ldm % 4 == 0), "skip_mix_bwd: D and the row strides must be multiples of 4 floats");
  const int n_rows = segs.ptr[T];
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_CUDA(cudaMemsetAsync(d_alpha, 0, (size_t)T * sizeof(float), wsi_stream(stream)));
  int blocks = (n_rows + 7) / 8;
  { const int sms = wsi_num_sms(); if (sms <= 0) return WSI_ERR_CUDA; if (blocks > sms * 16) blocks = sms * 16; }
  skip_mix_bwd_kernel<<<blocks, 256, 0, wsi_stream(stream)>>>(dout, ldd, out, ldo, x, ldx, drop_mask, ldm, skip, row_gate, segs, D,
                                                             d_lin, ldl, d_x, lddx, d_alpha);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
