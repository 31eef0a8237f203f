// Heterogeneous edge attention forward (kernel K2): one launch per layer for ALL relations.
//
// Replaces, per relation, the reference's apply_edges(fn.v_dot_u) + score scaling + dgl edge_softmax +
// update_all(u_mul_e, sum) and the cross-relation mean of multi_update_all
// (models/HEATNet4.py:103-119 == models/HEATNet2.py:78-94; models/HGT.py:95-106).
//
// Work item = an edge range inside one dst row of the relation-grouped CSR (HEAT) or inside one (dst, relation)
// segment (HGT).  Without a work list every row / segment is one item; with one (wsi_hetero_attn_work_fwd) rows with
// many in-edges (k-NN hubs: in-degree is heavy tailed although out-degree is fixed) are cut into chunks of one
// segment each, whose online-softmax partials (max, sum, unnormalised accumulator) are combined by attn_merge_kernel,
// so that no warp walks a 100+ edge row alone while the other SMs idle.
// One warp per work item: the 32 lanes span the D feature columns, the source rows K[src], V[src] are
// gathered with coalesced 16-byte loads (512 B per warp instruction), GROUP edges at a time (all K rows of the group in
// flight, then all V rows), the per-(segment, head) softmax is computed online (running max / running sum, rescaled
// accumulator) so every gathered byte is touched once.
//
// HBM-bound: algorithmic bytes per edge = 2*D*4 (K and V rows) + 9 (src id, sim, relation slot);
// per dst row = 2*D*4 (q in, agg out) + 8 (rowptr, 1/R).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr unsigned FULL = 0xffffffffu;

// sum over the G = 32 / H lanes of a head (G a power of two): five warp-uniform predicated steps - as a run-time loop this
// was 8 instructions per step and edge (loop counter, divergence check, branch) around one SHFL + one FADD
__device__ __forceinline__ float head_reduce(float d, int G) {
  if (G >= 32) d += __shfl_xor_sync(0xffffffffu, d, 16);
  if (G >= 16) d += __shfl_xor_sync(0xffffffffu, d, 8);
  if (G >= 8) d += __shfl_xor_sync(0xffffffffu, d, 4);
  if (G >= 4) d += __shfl_xor_sync(0xffffffffu, d, 2);
  if (G >= 2) d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}


enum { MODE_HEAT = 0, MODE_HGT_SEG = 1 };

struct AttnArgs {
  const float* K; int64_t ldk;
  const float* V; int64_t ldv;
  const float* Q; int64_t ldq;
  const int* rowptr;          // [n_items + 1] edge range of the work item
  const int* e_src;
  const float* e_sim;         // HEAT
  const uint8_t* e_rel;       // HEAT: relation slot per edge (segment boundaries inside a row)
  const float* inv_r;         // HEAT: 1/R_t per row (0 => passthrough row, written as 0)
  const float* e_w; const float* e_b;   // HEAT: e_linear scalars (device)
  const int* seg_rel;         // HGT: model relation id per segment
  const float* rel_pri;       // HGT: [R_model, H]
  int n_items;
  int D, H, dk;
  float inv_sqrt_dk;
  float* out; int64_t ldo;
  float* attn;                // optional [E, H] normalised attention (for backward)
  const int4* items;          // optional work list [n_items]: (row, e_beg, e_end, slot); slot < 0: the whole row
  float* part_ms;             // [P, 2, 32] per-lane (max, sum) of partial slot p
  float* part_acc;            // [P, D] unnormalised accumulator of partial slot p (physical column order)
  __nv_bfloat16* out_split;   // optional operand-form copy of `out` (16-bit elements): A operand of the a_linear GEMM
  int64_t split_lo;           // > 0: WSI_OPF_BF16X3, element offset of the lo half (n_rows * D); 0: WSI_OPF_F16; < 0: WSI_OPF_BF16
  // fused merge of the split rows (TMA kernel): split index of every partial, per-split-row arrival counter
  const int* part_split; int* split_cnt;
  int dbg;                    // development only (env WSI_ATTN_DEBUG): 1 = gather only 64 distinct rows, 2 = no bulk copies, 3 = no math
  int* sched;                 // optional int32 [2], zero before the first launch: dynamic work queue (next item | warps done)
  const int* split_row; const int* split_ptr; const int* part_rel;
  int kv_dtype;               // storage of K / V: 0 = fp32 (every kernel), 1 = fp16, 2 = bf16 (natural-order kernel only)
  int q_dtype;                // storage of Q: 0 = fp32, 1 = fp16, 2 = bf16 (lane-grouped kernels only; ldq counts elements)
  int64_t n_src_rows;         // rows of K / V that can be gathered (the footprint that decides register vs TMA-ring kernel); 0 = unknown
};

__device__ __forceinline__ void st_split4(__nv_bfloat16* dst, int64_t lo_off, float4 v) {
  if (lo_off == 0) {                                     // single fp16 operand, clamped to the finite range
    const float lim = 65504.f;
    __half2 a = __floats2half2_rn(fminf(fmaxf(v.x, -lim), lim), fminf(fmaxf(v.y, -lim), lim));
    __half2 b = __floats2half2_rn(fminf(fmaxf(v.z, -lim), lim), fminf(fmaxf(v.w, -lim), lim));
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst) = r;
    return;
  }
  if (lo_off < 0) {                                      // single bf16 operand
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a); r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst) = r;
    return;
  }
  __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
  wsi_split_bf16(v.x, h0, l0); wsi_split_bf16(v.y, h1, l1); wsi_split_bf16(v.z, h2, l2); wsi_split_bf16(v.w, h3, l3);
  __nv_bfloat162 hv[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
  __nv_bfloat162 lv[2] = {__halves2bfloat162(l0, l1), __halves2bfloat162(l2, l3)};
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<uint2*>(hv);
  *reinterpret_cast<uint2*>(dst + lo_off) = *reinterpret_cast<uint2*>(lv);
}

// 4 consecutive K / V elements of a storage type (fp32: 16 B, fp16 / bf16: 8 B) -> fp32
__device__ __forceinline__ float4 cvt4(uint2 r, const __half*) {
  const __half2 a = *reinterpret_cast<const __half2*>(&r.x), b = *reinterpret_cast<const __half2*>(&r.y);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 cvt4(uint2 r, const __nv_bfloat16*) {
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&r.x), b = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ float4 ldkv4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldkv4(const __half* p) { return cvt4(__ldg(reinterpret_cast<const uint2*>(p)), p); }
__device__ __forceinline__ float4 ldkv4(const __nv_bfloat16* p) { return cvt4(__ldg(reinterpret_cast<const uint2*>(p)), p); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 lds4(const __half* p) { return cvt4(*reinterpret_cast<const uint2*>(p), p); }
__device__ __forceinline__ float4 lds4(const __nv_bfloat16* p) { return cvt4(*reinterpret_cast<const uint2*>(p), p); }

// this lane's i-th 4-element slot of Q row `row` (fp32, or a 16-bit storage form converted on load)
__device__ __forceinline__ float4 ldq4(const AttnArgs& a, int64_t row, int col) {
  if (a.q_dtype == 0) return __ldg(reinterpret_cast<const float4*>(a.Q + row * a.ldq + col));
  if (a.q_dtype == 1) return ldkv4(reinterpret_cast<const __half*>(a.Q) + row * a.ldq + col);
  return ldkv4(reinterpret_cast<const __nv_bfloat16*>(a.Q) + row * a.ldq + col);
}

struct MergeArgs;
template <int NV> __device__ __noinline__ void merge_row(const MergeArgs& a, int h, int lane);
template <int NV> __device__ __forceinline__ void vec_fused_merge(const AttnArgs& a, int slot, int lane);

// ------------------------------------------------------------------------------------------------
// Lane-grouped fast path: D = 128*NV, H | 32.  Lane l owns float4 slots {i*32 + l}, all of head l / G
// (G = 32/H lanes per head) thanks to the head_perm column order (wsi_head_perm).
// GROUP = edges whose K (then V) rows are in flight together (bounded by the register file: GROUP * NV float4)
// KVT = storage type of K / V (float, or __half / __nv_bfloat16: half the gathered bytes, fp32 arithmetic; same
// lane-grouped column order, a lane's 4-element slot is then an 8-byte load); ldk / ldv count elements.
template <int NV, int MODE, int GROUP, int MINB, typename KVT>
__global__ void __launch_bounds__(WARPS * 32, MINB) attn_fwd_vec_kernel(AttnArgs a) {
  const int lane = threadIdx.x & 31;
  const int G = 32 / a.H;
  const int head = lane / G;
  const int n_warps = gridDim.x * WARPS;
  float ew = 0.f, eb = 0.f;
  if (MODE == MODE_HEAT) { ew = __ldg(a.e_w); eb = __ldg(a.e_b); }

  wsi_pdl_trigger();
  bool upstream_ready = false;      // PDL: the item descriptor and 1/R (plan data) are fetched before the previous kernel is waited for
  for (int item = blockIdx.x * WARPS + (threadIdx.x >> 5); item < a.n_items; item += n_warps) {
    int row = item, beg, end, slot = -1;
    if (a.items) {
      const int4 it = __ldg(a.items + item);
      row = it.x; beg = it.y; end = it.z; slot = it.w;
    } else {
      beg = __ldg(a.rowptr + item); end = __ldg(a.rowptr + item + 1);
    }
    float4 out[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float invr = 1.f, seg_scale = 0.f;
    if (MODE == MODE_HEAT) invr = __ldg(a.inv_r + row);
    if (!upstream_ready) { wsi_pdl_wait(); upstream_ready = true; }
    if (MODE != MODE_HEAT) seg_scale = __ldg(a.rel_pri + (int64_t)__ldg(a.seg_rel + row) * a.H + head) * a.inv_sqrt_dk;

    float m = -INFINITY, ssum = 0.f;
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    if (invr != 0.f && end > beg) {
      float4 q[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) q[i] = ldq4(a, row, (i * 32 + lane) * 4);
      int cur_rel = -1, seg_beg = beg;

      for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_src = 0, my_rel = 0;
        float my_sim = 0.f;
        if (lane < n) {
          my_src = __ldg(a.e_src + base + lane);
          if (MODE == MODE_HEAT) { my_sim = __ldg(a.e_sim + base + lane); my_rel = __ldg(a.e_rel + base + lane); }
        }
        int j = 0;
        while (j < n) {
          int g = min(GROUP, n - j);
          if (MODE == MODE_HEAT) {
            const int rel = __shfl_sync(FULL, my_rel, j);
            // edges j.. of the same relation (segments are contiguous): run length, capped at GROUP
            const unsigned same = __ballot_sync(FULL, lane >= j && lane < n && my_rel == rel) >> j;
            g = min(g, same == FULL ? 32 : __ffs(~same) - 1);   // (__ffs(0) == 0)
            if (rel != cur_rel) {                       // warp-uniform: close the running segment
              if (cur_rel >= 0) {
                const float inv = 1.f / ssum;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                  out[i].x = fmaf(acc[i].x, inv, out[i].x); out[i].y = fmaf(acc[i].y, inv, out[i].y);
                  out[i].z = fmaf(acc[i].z, inv, out[i].z); out[i].w = fmaf(acc[i].w, inv, out[i].w);
                  acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (a.attn && lane % G == 0)
                  for (int e = seg_beg; e < base + j; ++e) {
                    float* p = a.attn + (int64_t)e * a.H + head;
                    *p = __expf(*p - m) * inv;
                  }
              }
              m = -INFINITY; ssum = 0.f; cur_rel = rel; seg_beg = base + j;
            }
          }
          // ---- K rows of the group, scores
          float4 buf[GROUP][NV];
          float sc[GROUP];
#pragma unroll
          for (int u = 0; u < GROUP; ++u) {
            if (u < g) {
              const int src = __shfl_sync(FULL, my_src, j + u);
              const KVT* kr = reinterpret_cast<const KVT*>(a.K) + (int64_t)src * a.ldk;
#pragma unroll
              for (int i = 0; i < NV; ++i) buf[u][i] = ldkv4(kr + (i * 32 + lane) * 4);
            }
          }
#pragma unroll
          for (int u = 0; u < GROUP; ++u) {
            sc[u] = -INFINITY;
            if (u < g) {
              float d = 0.f;
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                d = fmaf(q[i].x, buf[u][i].x, d); d = fmaf(q[i].y, buf[u][i].y, d);
                d = fmaf(q[i].z, buf[u][i].z, d); d = fmaf(q[i].w, buf[u][i].w, d);
              }
              d = head_reduce(d, G);
              float scale = seg_scale;
              if (MODE == MODE_HEAT) scale = fmaf(ew, __shfl_sync(FULL, my_sim, j + u), eb) * a.inv_sqrt_dk;
              sc[u] = d * scale;
              if (a.attn && lane % G == 0) a.attn[(int64_t)(base + j + u) * a.H + head] = sc[u];
            }
          }
          // ---- V rows of the group (issued before the exponentials)
#pragma unroll
          for (int u = 0; u < GROUP; ++u) {
            if (u < g) {
              const int src = __shfl_sync(FULL, my_src, j + u);
              const KVT* vr = reinterpret_cast<const KVT*>(a.V) + (int64_t)src * a.ldv;
#pragma unroll
              for (int i = 0; i < NV; ++i) buf[u][i] = ldkv4(vr + (i * 32 + lane) * 4);
            }
          }
          float mn = m;
#pragma unroll
          for (int u = 0; u < GROUP; ++u) mn = fmaxf(mn, sc[u]);
          const float corr = __expf(m - mn);            // m = -inf on the first group -> 0
          float p[GROUP], psum = 0.f;
#pragma unroll
          for (int u = 0; u < GROUP; ++u) { p[u] = __expf(sc[u] - mn); psum += p[u]; }   // exp(-inf) = 0 for u >= g
          ssum = fmaf(ssum, corr, psum);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            acc[i].x *= corr; acc[i].y *= corr; acc[i].z *= corr; acc[i].w *= corr;
          }
#pragma unroll
          for (int u = 0; u < GROUP; ++u) {
            if (u < g) {
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                acc[i].x = fmaf(p[u], buf[u][i].x, acc[i].x); acc[i].y = fmaf(p[u], buf[u][i].y, acc[i].y);
                acc[i].z = fmaf(p[u], buf[u][i].z, acc[i].z); acc[i].w = fmaf(p[u], buf[u][i].w, acc[i].w);
              }
            }
          }
          m = mn;
          j += g;
        }
      }
      if (slot < 0) {                                   // close the last segment
        const float inv = 1.f / ssum;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          out[i].x = fmaf(acc[i].x, inv, out[i].x) * invr; out[i].y = fmaf(acc[i].y, inv, out[i].y) * invr;
          out[i].z = fmaf(acc[i].z, inv, out[i].z) * invr; out[i].w = fmaf(acc[i].w, inv, out[i].w) * invr;
        }
        if (a.attn && lane % G == 0)
          for (int e = seg_beg; e < end; ++e) {
            float* p = a.attn + (int64_t)e * a.H + head;
            *p = __expf(*p - m) * inv;
          }
      }
    }
    if (slot < 0) {
      if (a.out) {
        float* o = a.out + (int64_t)row * a.ldo;
#pragma unroll
        for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = out[i];
      }
      if (a.out_split) {
        __nv_bfloat16* o = a.out_split + (int64_t)row * a.D;
#pragma unroll
        for (int i = 0; i < NV; ++i) st_split4(o + (i * 32 + lane) * 4, a.split_lo, out[i]);
      }
    } else {                                            // partial of one segment chunk: (m, sum, unnormalised acc)
      a.part_ms[(int64_t)slot * 64 + lane] = m;
      a.part_ms[(int64_t)slot * 64 + 32 + lane] = ssum;
      float* o = a.part_acc + (int64_t)slot * a.D;
#pragma unroll
      for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = acc[i];
      if (a.split_cnt) vec_fused_merge<NV>(a, slot, lane);
    }
  }
  if (!upstream_ready) wsi_pdl_wait();                  // (a warp without an item)
}

// Combine the chunk partials of the split rows (HEAT) / split segments (HGT): one warp per split row.
//   split_row[h] = output row, split_ptr[h..h+1] = its partial slots (in edge order), part_rel[p] = relation slot
//   of partial p (a new value closes the running segment).  out[row] = inv_r[row] * sum_segments acc / sum.
// Runs either fused into the TMA kernel (the warp that finishes the LAST chunk of a row merges it: per-row arrival
// counter split_cnt, self-resetting) or as its own launch after the register-path kernel.
struct MergeArgs {
  const int* split_row; const int* split_ptr; const int* part_rel;
  const float* part_ms; const float* part_acc; const float* inv_r;
  int n_split, D;
  float* out; int64_t ldo;
  __nv_bfloat16* out_split; int64_t split_lo;
};

template <int NV>
__device__ __noinline__ void merge_row(const MergeArgs& a, int h, int lane) {
  constexpr int MB = 4;                                 // partials whose loads are in flight together
  const int row = __ldg(a.split_row + h);
  const int pb = __ldg(a.split_ptr + h), pe = __ldg(a.split_ptr + h + 1);
  const float invr = a.inv_r ? __ldg(a.inv_r + row) : 1.f;
  float4 out[NV], acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) { out[i] = make_float4(0.f, 0.f, 0.f, 0.f); acc[i] = out[i]; }
  float m = -INFINITY, ssum = 0.f;
  int cur_rel = -1;
  for (int p0 = pb; p0 < pe; p0 += MB) {
    const int nb = min(MB, pe - p0);
    int rl[MB];
    float mp[MB], sp[MB];
    float4 x[MB][NV];
#pragma unroll
    for (int u = 0; u < MB; ++u) {
      if (u < nb) {                                     // partials were written by this kernel: L2-coherent loads
        const int p = p0 + u;
        rl[u] = __ldg(a.part_rel + p);
        mp[u] = __ldcg(a.part_ms + (int64_t)p * 64 + lane);
        sp[u] = __ldcg(a.part_ms + (int64_t)p * 64 + 32 + lane);
        const float* pa = a.part_acc + (int64_t)p * a.D;
#pragma unroll
        for (int i = 0; i < NV; ++i) x[u][i] = __ldcg(reinterpret_cast<const float4*>(pa + (i * 32 + lane) * 4));
      }
    }
#pragma unroll
    for (int u = 0; u < MB; ++u) {
      if (u < nb) {
        if (rl[u] != cur_rel) {
          if (cur_rel >= 0 && ssum > 0.f) {
            const float inv = 1.f / ssum;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              out[i].x = fmaf(acc[i].x, inv, out[i].x); out[i].y = fmaf(acc[i].y, inv, out[i].y);
              out[i].z = fmaf(acc[i].z, inv, out[i].z); out[i].w = fmaf(acc[i].w, inv, out[i].w);
            }
          }
#pragma unroll
          for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          m = -INFINITY; ssum = 0.f; cur_rel = rl[u];
        }
        if (sp[u] > 0.f) {                              // (an untouched partial has sum 0: passthrough row)
          const float mn = fmaxf(m, mp[u]);
          const float c0 = __expf(m - mn), c1 = __expf(mp[u] - mn);
          ssum = fmaf(ssum, c0, sp[u] * c1);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            acc[i].x = fmaf(acc[i].x, c0, x[u][i].x * c1); acc[i].y = fmaf(acc[i].y, c0, x[u][i].y * c1);
            acc[i].z = fmaf(acc[i].z, c0, x[u][i].z * c1); acc[i].w = fmaf(acc[i].w, c0, x[u][i].w * c1);
          }
          m = mn;
        }
      }
    }
  }
  if (cur_rel >= 0 && ssum > 0.f) {
    const float inv = 1.f / ssum;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      out[i].x = fmaf(acc[i].x, inv, out[i].x); out[i].y = fmaf(acc[i].y, inv, out[i].y);
      out[i].z = fmaf(acc[i].z, inv, out[i].z); out[i].w = fmaf(acc[i].w, inv, out[i].w);
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 r = make_float4(out[i].x * invr, out[i].y * invr, out[i].z * invr, out[i].w * invr);
    if (a.out) *reinterpret_cast<float4*>(a.out + (int64_t)row * a.ldo + (i * 32 + lane) * 4) = r;
    if (a.out_split) st_split4(a.out_split + (int64_t)row * a.D + (i * 32 + lane) * 4, a.split_lo, r);
  }
}

// Fused merge: the warp that writes the LAST chunk partial of a split row combines the row (per-row arrival counter,
// self-resetting so the next launch starts from zero).
template <int NV>
__device__ __forceinline__ void vec_fused_merge(const AttnArgs& a, int slot, int lane) {
  const int h = __ldg(a.part_split + slot);
  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) {
    const int n_chunks = __ldg(a.split_ptr + h + 1) - __ldg(a.split_ptr + h);
    last = atomicAdd(a.split_cnt + h, 1) == n_chunks - 1;
    if (last) a.split_cnt[h] = 0;
  }
  last = __shfl_sync(FULL, last, 0);
  if (last) {
    __threadfence();
    MergeArgs mg;
    mg.split_row = a.split_row; mg.split_ptr = a.split_ptr; mg.part_rel = a.part_rel;
    mg.part_ms = a.part_ms; mg.part_acc = a.part_acc; mg.inv_r = a.inv_r; mg.n_split = 0; mg.D = a.D;
    mg.out = a.out; mg.ldo = a.ldo; mg.out_split = a.out_split; mg.split_lo = a.split_lo;
    merge_row<NV>(mg, h, lane);
  }
}

template <int NV>
__global__ void __launch_bounds__(WARPS * 32) attn_merge_kernel(MergeArgs a) {
  const int h = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (h < a.n_split) merge_row<NV>(a, h, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged variant of the fast path (chosen per launch for K | V footprints beyond L2): every warp owns a ring of
// shared-memory slots, one slot = the K row and the V row of one edge (2 * D * element size), filled by cp.async.bulk
// (the TMA engine's 1-D bulk copy, SASS UBLKCP) and signalled through one mbarrier per slot.  The edges of a work item
// (up to the ring depth) are in flight at once while the registers only hold q / acc / out; the copies of a batch are
// issued by the lanes in one predicated pass (lane u fills slot rs + u).  Six resident blocks of four warps with
// ~8-10 KB of rows in flight per warp: ~190 KB per SM.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int TMA_WARPS = 4;
constexpr int BMAX = 2;      // edges per batch of the TMA kernel (4 measured no faster: 408 vs 410 us on config 3, 506 vs 498 us on
                             // config 4 - and costs 16 more registers, i.e. the sixth resident block)

template <int NV, int MODE, typename KVT>
__global__ void __launch_bounds__(TMA_WARPS * 32, NV <= 4 ? 6 : 2) attn_fwd_tma_kernel(AttnArgs a, int ring) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int ROW_BYTES = NV * 128 * (int)sizeof(KVT), SLOT_BYTES = 2 * ROW_BYTES;
  const KVT* const Kp = reinterpret_cast<const KVT*>(a.K);
  const KVT* const Vp = reinterpret_cast<const KVT*>(a.V);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = 32 / a.H;
  const int head = lane / G;
  const int n_warps = gridDim.x * TMA_WARPS;
  uint8_t* my_slots = smem + (size_t)warp * ring * SLOT_BYTES;
  const uint32_t slots_u32 = smem_u32(my_slots);
  const uint32_t bars_u32 = smem_u32(smem + (size_t)TMA_WARPS * ring * SLOT_BYTES) + warp * ring * 8;
  if (lane == 0) {
    for (int s = 0; s < ring; ++s) mbar_init(bars_u32 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float ew = 0.f, eb = 0.f;
  if (MODE == MODE_HEAT) { ew = __ldg(a.e_w); eb = __ldg(a.e_b); }
  const bool kv_adjacent = Vp == Kp + a.D && a.ldk == a.ldv;
  int rs = 0;                 // ring slot of the next edge to consume
  uint32_t rpar = 0;          // its mbarrier phase parity

  // Work distribution: with a.sched the warps pull items from a device-side queue (the item list is sorted largest
  // first, so this is LPT scheduling and no warp is left with a long tail); the next index is fetched one item ahead
  // so the atomic's latency hides behind the current item.  Without it: static round-robin.
  int item = blockIdx.x * TMA_WARPS + warp;
  int next_item = 0;
  if (a.sched) {
    if (lane == 0) { item = atomicAdd(a.sched, 1); }
    item = __shfl_sync(FULL, item, 0);
  }
  while (item < a.n_items) {
    if (a.sched) { if (lane == 0) next_item = atomicAdd(a.sched, 1); }
    else next_item = item + n_warps;
    int row = item, beg, end, slot = -1;
    if (a.items) {
      const int4 it = __ldg(a.items + item);
      row = it.x; beg = it.y; end = it.z; slot = it.w;
    } else {
      beg = __ldg(a.rowptr + item); end = __ldg(a.rowptr + item + 1);
    }
    float4 out[NV], acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { out[i] = make_float4(0.f, 0.f, 0.f, 0.f); acc[i] = out[i]; }
    float invr = 1.f, seg_scale = 0.f;
    if (MODE == MODE_HEAT) invr = __ldg(a.inv_r + row);
    else seg_scale = __ldg(a.rel_pri + (int64_t)__ldg(a.seg_rel + row) * a.H + head) * a.inv_sqrt_dk;
    float m = -INFINITY, ssum = 0.f;

    if (invr != 0.f && end > beg) {
      int cur_rel = -1;
      // the first window's edge ids first: their K/V rows are requested before q is even loaded
      float4 q[NV];
      bool have_q = false;
      for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_src = 0, my_rel = 0;
        float my_sim = 0.f;
        if (lane < n) {
          my_src = __ldg(a.e_src + base + lane);
          if (a.dbg == 1) my_src &= 63;
          if (MODE == MODE_HEAT) { my_sim = __ldg(a.e_sim + base + lane); my_rel = __ldg(a.e_rel + base + lane); }
        }
        const int pre = min(n, ring);
        // every lane issues the copy of ITS edge (lane j < pre holds edge j): one predicated pass instead of a serial
        // loop through lane 0
        if (lane < pre && a.dbg != 2) {
          int s = rs + lane;
          if (s >= ring) s -= ring;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // (the slot was last read through the generic proxy)
          const uint32_t bar = bars_u32 + 8 * s, dst = slots_u32 + s * SLOT_BYTES;
          mbar_expect_tx(bar, SLOT_BYTES);
          if (kv_adjacent) {                            // K|V of a node are one contiguous 2 * D * 4 byte run
            bulk_g2s(dst, Kp + (int64_t)my_src * a.ldk, SLOT_BYTES, bar);
          } else {
            bulk_g2s(dst, Kp + (int64_t)my_src * a.ldk, ROW_BYTES, bar);
            bulk_g2s(dst + ROW_BYTES, Vp + (int64_t)my_src * a.ldv, ROW_BYTES, bar);
          }
        }
        if (!have_q) {
#pragma unroll
          for (int i = 0; i < NV; ++i) q[i] = ldq4(a, row, (i * 32 + lane) * 4);
          have_q = true;
        }
        // edges are consumed in batches of up to BMAX (same relation): all scores first (independent dot products),
        // one rescale of the accumulator per batch, then the weighted V rows
        int j = 0;
        while (j < n) {
          int g = min(min(BMAX, ring), n - j);
          if (MODE == MODE_HEAT) {
            const int rel = __shfl_sync(FULL, my_rel, j);
            const unsigned same = __ballot_sync(FULL, lane >= j && lane < n && my_rel == rel) >> j;
            g = min(g, same == FULL ? 32 : __ffs(~same) - 1);        // (__ffs(0) == 0)
            if (rel != cur_rel) {                       // warp-uniform: close the running segment
              if (cur_rel >= 0) {
                const float inv = 1.f / ssum;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                  out[i].x = fmaf(acc[i].x, inv, out[i].x); out[i].y = fmaf(acc[i].y, inv, out[i].y);
                  out[i].z = fmaf(acc[i].z, inv, out[i].z); out[i].w = fmaf(acc[i].w, inv, out[i].w);
                  acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
              m = -INFINITY; ssum = 0.f; cur_rel = rel;
            }
          }
          float sc[BMAX];
          const KVT* slot_ptr[BMAX];
#pragma unroll
          for (int u = 0; u < BMAX; ++u) {
            sc[u] = -INFINITY;
            slot_ptr[u] = nullptr;
            if (u < g) {
              int su = rs + u;
              uint32_t pu = rpar;
              if (su >= ring) { su -= ring; pu ^= 1; }
              if (a.dbg != 2) mbar_wait(bars_u32 + 8 * su, pu);
              const KVT* ks = reinterpret_cast<const KVT*>(my_slots + (size_t)su * SLOT_BYTES);
              slot_ptr[u] = ks;
              float d0 = 0.f, d1 = 0.f;
              if (a.dbg != 3)
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                const float4 kk = lds4(ks + (i * 32 + lane) * 4);
                d0 = fmaf(q[i].x, kk.x, d0); d1 = fmaf(q[i].y, kk.y, d1);
                d0 = fmaf(q[i].z, kk.z, d0); d1 = fmaf(q[i].w, kk.w, d1);
              }
              sc[u] = d0 + d1;
            }
          }
#pragma unroll
          for (int u = 0; u < BMAX; ++u) {
            if (u < g) {
              float d = sc[u];
              d = head_reduce(d, G);
              float scale = seg_scale;
              if (MODE == MODE_HEAT) scale = fmaf(ew, __shfl_sync(FULL, my_sim, j + u), eb) * a.inv_sqrt_dk;
              sc[u] = d * scale;
            }
          }
          float mn = m;
#pragma unroll
          for (int u = 0; u < BMAX; ++u) mn = fmaxf(mn, sc[u]);
          const float corr = __expf(m - mn);            // m = -inf on the first batch -> 0
          float p[BMAX], psum = 0.f;
#pragma unroll
          for (int u = 0; u < BMAX; ++u) { p[u] = __expf(sc[u] - mn); psum += p[u]; }   // exp(-inf) = 0 for u >= g
          ssum = fmaf(ssum, corr, psum);
          m = mn;
#pragma unroll
          for (int i = 0; i < NV; ++i) { acc[i].x *= corr; acc[i].y *= corr; acc[i].z *= corr; acc[i].w *= corr; }
#pragma unroll
          for (int u = 0; u < BMAX; ++u) {
            if (u < g && a.dbg != 3) {
              const KVT* vs = slot_ptr[u] + NV * 128;
#pragma unroll
              for (int i = 0; i < NV; ++i) {
                const float4 vv = lds4(vs + (i * 32 + lane) * 4);
                acc[i].x = fmaf(p[u], vv.x, acc[i].x); acc[i].y = fmaf(p[u], vv.y, acc[i].y);
                acc[i].z = fmaf(p[u], vv.z, acc[i].z); acc[i].w = fmaf(p[u], vv.w, acc[i].w);
              }
            }
          }
          __syncwarp();                                 // every lane is done with these g slots
          {                                             // refill them with the edges `ring` positions ahead: lane u < g
            const int nxt = j + lane + ring;            // takes slot rs + u, all g copies issued in one predicated pass
            const int src = __shfl_sync(FULL, my_src, nxt & 31);
            if (lane < g && nxt < n && a.dbg != 2) {
              int s = rs + lane;
              if (s >= ring) s -= ring;
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              const uint32_t bar = bars_u32 + 8 * s, dst = slots_u32 + s * SLOT_BYTES;
              mbar_expect_tx(bar, SLOT_BYTES);
              if (kv_adjacent) {                        // K|V of a node are one contiguous 2 * D * 4 byte run
                bulk_g2s(dst, Kp + (int64_t)src * a.ldk, SLOT_BYTES, bar);
              } else {
                bulk_g2s(dst, Kp + (int64_t)src * a.ldk, ROW_BYTES, bar);
                bulk_g2s(dst + ROW_BYTES, Vp + (int64_t)src * a.ldv, ROW_BYTES, bar);
              }
            }
            rs += g;
            if (rs >= ring) { rs -= ring; rpar ^= 1; }
          }
          j += g;
        }
      }
      if (slot < 0) {                                   // close the last segment
        const float inv = 1.f / ssum;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          out[i].x = fmaf(acc[i].x, inv, out[i].x) * invr; out[i].y = fmaf(acc[i].y, inv, out[i].y) * invr;
          out[i].z = fmaf(acc[i].z, inv, out[i].z) * invr; out[i].w = fmaf(acc[i].w, inv, out[i].w) * invr;
        }
      }
    }
    if (slot < 0) {
      if (a.out) {
        float* o = a.out + (int64_t)row * a.ldo;
#pragma unroll
        for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = out[i];
      }
      if (a.out_split) {
        __nv_bfloat16* o = a.out_split + (int64_t)row * a.D;
#pragma unroll
        for (int i = 0; i < NV; ++i) st_split4(o + (i * 32 + lane) * 4, a.split_lo, out[i]);
      }
    } else {                                            // partial of one segment chunk: (m, sum, unnormalised acc)
      a.part_ms[(int64_t)slot * 64 + lane] = m;
      a.part_ms[(int64_t)slot * 64 + 32 + lane] = ssum;
      float* o = a.part_acc + (int64_t)slot * a.D;
#pragma unroll
      for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = acc[i];
      if (a.split_cnt) {                                // the warp that completes the row's last chunk merges it
        const int h = __ldg(a.part_split + slot);
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          const int n_chunks = __ldg(a.split_ptr + h + 1) - __ldg(a.split_ptr + h);
          last = atomicAdd(a.split_cnt + h, 1) == n_chunks - 1;
          if (last) a.split_cnt[h] = 0;                 // self-resetting: ready for the next launch
        }
        last = __shfl_sync(FULL, last, 0);
        if (last) {
          __threadfence();
          MergeArgs mg;
          mg.split_row = a.split_row; mg.split_ptr = a.split_ptr; mg.part_rel = a.part_rel;
          mg.part_ms = a.part_ms; mg.part_acc = a.part_acc; mg.inv_r = a.inv_r; mg.n_split = 0; mg.D = a.D;
          mg.out = a.out; mg.ldo = a.ldo; mg.out_split = a.out_split; mg.split_lo = a.split_lo;
          merge_row<NV>(mg, h, lane);
        }
      }
    }
    item = a.sched ? __shfl_sync(FULL, next_item, 0) : next_item;
  }
  if (a.sched && lane == 0) {                           // the last warp to leave re-arms the queue for the next launch
    __threadfence();
    if (atomicAdd(a.sched + 1, 1) == n_warps - 1) { a.sched[0] = 0; a.sched[1] = 0; }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic path: any D, H (natural column order).  One warp per work item, heads processed one after the
// other; lane l owns columns {l + 32 j} of the current head (MAXJ >= ceil(d_k / 32)).
// KVT = storage type of K / V (float, __half, __nv_bfloat16: the 16-bit forms are the bf16-storage configuration,
// BASELINE config 3 - gathered bytes halve, the arithmetic stays fp32); ldk / ldv count elements.
__device__ __forceinline__ float kv_ld(const float* p) { return __ldg(p); }
__device__ __forceinline__ float kv_ld(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ float kv_ld(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

template <int MAXJ, int MODE, typename KVT>
__global__ void __launch_bounds__(WARPS * 32) attn_fwd_generic_kernel(AttnArgs a) {
  const int lane = threadIdx.x & 31;
  const int n_warps = gridDim.x * WARPS;
  const int dk = a.dk;
  float ew = 0.f, eb = 0.f;
  if (MODE == MODE_HEAT) { ew = __ldg(a.e_w); eb = __ldg(a.e_b); }

  for (int item = blockIdx.x * WARPS + (threadIdx.x >> 5); item < a.n_items; item += n_warps) {
    const int beg = __ldg(a.rowptr + item), end = __ldg(a.rowptr + item + 1);
    float invr = 1.f;
    int model_rel = 0;
    if (MODE == MODE_HEAT) invr = __ldg(a.inv_r + item);
    else model_rel = __ldg(a.seg_rel + item);
    const bool live = invr != 0.f && end > beg;
    for (int h = 0; h < a.H; ++h) {
      float out[MAXJ];
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) out[j] = 0.f;
      if (live) {
        const float seg_scale = MODE == MODE_HEAT ? 0.f : __ldg(a.rel_pri + (int64_t)model_rel * a.H + h) * a.inv_sqrt_dk;
        float q[MAXJ], acc[MAXJ];
        const float* qr = a.Q + (int64_t)item * a.ldq + h * dk;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          int c = lane + 32 * j;
          q[j] = c < dk ? __ldg(qr + c) : 0.f;
          acc[j] = 0.f;
        }
        float m = -INFINITY, ssum = 0.f;
        int cur_rel = -1, seg_beg = beg;
        for (int e = beg; e < end; ++e) {
          const int src = __ldg(a.e_src + e);
          float scale = seg_scale;
          if (MODE == MODE_HEAT) {
            const int rel = __ldg(a.e_rel + e);
            if (rel != cur_rel) {
              if (cur_rel >= 0) {
                const float inv = 1.f / ssum;
#pragma unroll
                for (int j = 0; j < MAXJ; ++j) { out[j] = fmaf(acc[j], inv, out[j]); acc[j] = 0.f; }
                if (a.attn && lane == 0)
                  for (int e2 = seg_beg; e2 < e; ++e2) {
                    float* p = a.attn + (int64_t)e2 * a.H + h;
                    *p = __expf(*p - m) * inv;
                  }
              }
              m = -INFINITY; ssum = 0.f; cur_rel = rel; seg_beg = e;
            }
            scale = fmaf(ew, __ldg(a.e_sim + e), eb) * a.inv_sqrt_dk;
          }
          const KVT* kr = reinterpret_cast<const KVT*>(a.K) + (int64_t)src * a.ldk + h * dk;
          const KVT* vr = reinterpret_cast<const KVT*>(a.V) + (int64_t)src * a.ldv + h * dk;
          float vv[MAXJ];
          float d = 0.f;
#pragma unroll
          for (int j = 0; j < MAXJ; ++j) {
            int c = lane + 32 * j;
            float kk = c < dk ? kv_ld(kr + c) : 0.f;
            vv[j] = c < dk ? kv_ld(vr + c) : 0.f;
            d = fmaf(q[j], kk, d);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
          const float s = d * scale;
          if (a.attn && lane == 0) a.attn[(int64_t)e * a.H + h] = s;
          const float mn = fmaxf(m, s);
          const float corr = __expf(m - mn);
          const float p = __expf(s - mn);
          ssum = fmaf(ssum, corr, p);
#pragma unroll
          for (int j = 0; j < MAXJ; ++j) acc[j] = fmaf(acc[j], corr, p * vv[j]);
          m = mn;
        }
        const float inv = 1.f / ssum;
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) out[j] = fmaf(acc[j], inv, out[j]) * invr;
        if (a.attn && lane == 0)
          for (int e2 = seg_beg; e2 < end; ++e2) {
            float* p = a.attn + (int64_t)e2 * a.H + h;
            *p = __expf(*p - m) * inv;
          }
      }
      float* o = a.out + (int64_t)item * a.ldo + h * dk;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        int c = lane + 32 * j;
        if (c < dk) o[c] = out[j];
      }
    }
  }
}

bool vec_ok(int D, int H) { return D % 128 == 0 && D <= 1024 && H >= 1 && H <= 32 && (H & (H - 1)) == 0; }

// *fused_merge (optional): set to true when the launched kernel merges the split rows itself
template <int MODE>
int launch(const AttnArgs& a_in, int head_perm, cudaStream_t stream, bool* fused_merge = nullptr) {
  AttnArgs a = a_in;
  if (fused_merge) *fused_merge = false;
  if (a.n_items == 0) return WSI_OK;
  int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  int blocks = (a.n_items + WARPS - 1) / WARPS;
  // one work item per warp: the hardware block scheduler deals the (largest-first) items out as blocks retire, which
  // balances better than a persistent grid-stride loop; WSI_ATTN_CAP (development knob) = blocks per SM of a persistent grid
  const WsiDev& dev = *wsi_dev();
  if (dev.attn_cap > 0 && blocks > sms * dev.attn_cap) blocks = sms * dev.attn_cap;
  if (head_perm) {
    if (!vec_ok(a.D, a.H)) {
      wsi_set_error("hetero_attn: head_perm layout needs D %% 128 == 0, D <= 1024, H a power of two <= 32 (D=%d H=%d)", a.D, a.H);
      return WSI_ERR_UNSUPPORTED;
    }
    // Kernel choice per launch (round-2 sweep, profiles/r2_attn_crossover.txt): while K|V fit in L2 (config 2: 33 MB)
    // the register-path kernel wins (30.7 vs 43.2 us) - the gathers are L2 hits and many light warps hide their latency
    // best; once the gathered rows come from DRAM (>= ~16-20k rows at D = 512; config 4: 410 MB, the packed training
    // batches: 655 MB) the TMA bulk-copy ring wins by 1.2-1.3x (100k nodes, k = 8: 0.537 vs 0.707 ms = 6.9 TB/s on the
    // SURVEY 8(d) byte model) because it keeps ~12 KB per warp in flight without spending registers on them.  The
    // item-pipelined variant of round 1 lost at every size and no longer ships.  Knob attn_kernel: 1 / 2 force one.
    const int es = a.kv_dtype == 0 ? 4 : 2;                     // bytes per stored K / V element
    const int64_t kv_bytes = a.n_src_rows * 2 * (int64_t)a.D * es;
    const bool want_ring = dev.attn_kernel == 2 || (dev.attn_kernel == 0 && kv_bytes >= (80ll << 20));
    if (!a.attn && (a.ldk % 8 == 0) && (a.ldv % 8 == 0) && want_ring) {      // (bulk copies need 16 B aligned rows)
      const int slot_bytes = 2 * a.D * es;
      // The kernel is bound by latency per warp, not by bytes in flight (round-2 sweep on the HGT segment graph of
      // config 3: time ~ 1 / resident blocks from 1 to 4 blocks per SM, flat in the ring depth from 2 to 4 slots), so the
      // ring is sized for FIVE to SIX resident blocks of 4 warps (80 registers, <= 40 KB of shared memory each): ~10 KB of
      // K/V rows in flight per warp (config 3, bf16 rows: 555 us at 4 blocks -> 462 us at 5 -> 396 us at 6 with the
      // lane-parallel copy issue; config 4, fp32: 530 -> 498 us)
      int ring = 9984 / slot_bytes;
      ring = ring < 2 ? 2 : (ring > 8 ? 8 : ring);
      if (dev.attn_ring > 0) ring = dev.attn_ring;                          // development knob
      if (ring < 1) ring = 1;
      const int smem = TMA_WARPS * ring * slot_bytes + TMA_WARPS * ring * 8;
      int tb = (a.n_items + TMA_WARPS - 1) / TMA_WARPS;
      const int per_sm = 200 * 1024 / (smem + 1024) < 1 ? 1 : 200 * 1024 / (smem + 1024);
      if (tb > sms * per_sm) tb = sms * per_sm;
      if (dev.attn_blocks > 0 && tb > sms * dev.attn_blocks) tb = sms * dev.attn_blocks;   // development knob
#define RING(NV, KVT) { \
        static std::once_flag once; \
        static cudaError_t aerr = cudaSuccess; \
        std::call_once(once, [] { \
          aerr = cudaFuncSetAttribute(attn_fwd_tma_kernel<NV, MODE, KVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }); \
        WSI_CHECK_CUDA(aerr); \
        attn_fwd_tma_kernel<NV, MODE, KVT><<<tb, TMA_WARPS * 32, smem, stream>>>(a, ring); }
#define CASE(NV) case NV: \
        if (a.kv_dtype == 1) RING(NV, __half) \
        else if (a.kv_dtype == 2) RING(NV, __nv_bfloat16) \
        else RING(NV, float) \
        break;
      switch (a.D / 128) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
      }
#undef CASE
#undef RING
      WSI_CHECK_LAUNCH();
      if (fused_merge) *fused_merge = a.split_cnt != nullptr;
      return WSI_OK;
    }
    // register-path kernel (default): one work item per warp, the block scheduler is the work queue; rows are gathered
    // with 16-byte loads straight into registers, GROUP edges in flight; split rows merged through arrival counters
    a.sched = nullptr;
    if (dev.attn_separate_merge) a.split_cnt = nullptr;                // development knob: merge as its own launch
    // GROUP edges in flight per warp x MINB resident blocks per SM: measured on config 2 (D = 512) the kernel is bound
    // by latency per warp, not by bytes in flight, so more (lighter) warps win: GROUP 2 at 4 blocks / SM (128
    // registers) beats GROUP 4 at 3, GROUP 1 at 5-6 and GROUP 2 at 5 (round-1 sweep; those variants no longer ship).
    switch (a.D / 128) {
#define VEC(NV, GRP, MINB, KVT) WSI_CHECK_CUDA(wsi_launch_pdl(attn_fwd_vec_kernel<NV, MODE, GRP, MINB, KVT>, dim3(blocks), dim3(WARPS * 32), 0, stream, a))
#define CASE(NV, GRP, MINB) case NV: \
      if (a.kv_dtype == 1) VEC(NV, GRP, MINB, __half); \
      else if (a.kv_dtype == 2) VEC(NV, GRP, MINB, __nv_bfloat16); \
      else VEC(NV, GRP, MINB, float); \
      break;
      CASE(1, 4, 4) CASE(2, 4, 4) CASE(3, 2, 4) CASE(4, 2, 4) CASE(5, 2, 2) CASE(6, 2, 2) CASE(7, 2, 2) CASE(8, 2, 2)
#undef CASE
#undef VEC
    }
    WSI_CHECK_LAUNCH();
    if (fused_merge) *fused_merge = a.split_cnt != nullptr;
    return WSI_OK;
  } else {
    int mj = (a.dk + 31) / 32;
    if (mj > 8) {
      wsi_set_error("hetero_attn: d_k=%d > 256 is not supported", a.dk);
      return WSI_ERR_UNSUPPORTED;
    }
#define GEN(KVT) \
    if (mj <= 1) attn_fwd_generic_kernel<1, MODE, KVT><<<blocks, WARPS * 32, 0, stream>>>(a); \
    else if (mj <= 2) attn_fwd_generic_kernel<2, MODE, KVT><<<blocks, WARPS * 32, 0, stream>>>(a); \
    else if (mj <= 4) attn_fwd_generic_kernel<4, MODE, KVT><<<blocks, WARPS * 32, 0, stream>>>(a); \
    else attn_fwd_generic_kernel<8, MODE, KVT><<<blocks, WARPS * 32, 0, stream>>>(a);
    if (a.kv_dtype == 1) { GEN(__half) } else if (a.kv_dtype == 2) { GEN(__nv_bfloat16) } else { GEN(float) }
#undef GEN
  }
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

int launch_merge(const MergeArgs& m, cudaStream_t stream) {
  if (m.n_split == 0) return WSI_OK;
  const int blocks = (m.n_split + WARPS - 1) / WARPS;
  switch (m.D / 128) {
#define CASE(NV) case NV: attn_merge_kernel<NV><<<blocks, WARPS * 32, 0, stream>>>(m); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
  }
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

}  // namespace

extern "C" int wsi_head_perm(int D, int H, int32_t* perm_host) {
  WSI_CHECK_ARG(perm_host, "head_perm: null output");
  if (!vec_ok(D, H) || D % H != 0) {
    wsi_set_error("head_perm: no lane-grouped layout for D=%d H=%d", D, H);
    return WSI_ERR_UNSUPPORTED;
  }
  const int G = 32 / H, dk = D / H, NV = D / 128;
  for (int i = 0; i < NV; ++i)
    for (int l = 0; l < 32; ++l)
      for (int c = 0; c < 4; ++c) {
        int head = l / G;
        int within = i * (G * 4) + (l % G) * 4 + c;
        perm_host[(i * 32 + l) * 4 + c] = head * dk + within;
      }
  return WSI_OK;
}

extern "C" int wsi_hetero_attn_fwd(const float* k, int64_t ldk, const float* v, int64_t ldv, const float* q,
                                   int64_t ldq, const int32_t* rowptr, const int32_t* e_src, const float* e_sim,
                                   const uint8_t* e_rel, const float* node_inv_r, const float* e_w,
                                   const float* e_b, int64_t n_rows, int D, int H, int head_perm, float* agg,
                                   int64_t ldo, float* attn_out, void* stream) {
  WSI_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31), "hetero_attn_fwd: bad n_rows");
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(k && v && q && rowptr && node_inv_r && e_w && e_b && agg, "hetero_attn_fwd: null pointer");
  WSI_CHECK_ARG(H >= 1 && D >= 1 && D % H == 0, "hetero_attn_fwd: D=%d is not a multiple of H=%d", D, H);
  WSI_CHECK_ARG(!head_perm || (ldk % 4 == 0 && ldv % 4 == 0 && ldq % 4 == 0 && ldo % 4 == 0),
                "hetero_attn_fwd: row strides must be multiples of 4 floats for the vector path");
  AttnArgs a{};
  a.K = k; a.ldk = ldk; a.V = v; a.ldv = ldv; a.Q = q; a.ldq = ldq;
  a.rowptr = rowptr; a.e_src = e_src; a.e_sim = e_sim; a.e_rel = e_rel; a.inv_r = node_inv_r;
  a.e_w = e_w; a.e_b = e_b; a.n_items = (int)n_rows; a.D = D; a.H = H; a.dk = D / H;
  a.inv_sqrt_dk = 1.0f / sqrtf((float)(D / H));
  a.out = agg; a.ldo = ldo; a.attn = attn_out;
  a.n_src_rows = n_rows;
  return launch<MODE_HEAT>(a, head_perm, wsi_stream(stream));
}

extern "C" int wsi_hetero_attn_work_fwd(const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_dtype, const void* q,
                                        int q_dtype, int64_t ldq, const int32_t* e_src, const float* e_sim, const uint8_t* e_rel,
                                        const float* node_inv_r, const float* e_w, const float* e_b, int64_t n_rows,
                                        int64_t n_src_rows, int D, int H, const int32_t* items, int64_t n_items,
                                        const int32_t* split_row, const int32_t* split_ptr, const int32_t* part_rel,
                                        const int32_t* part_split, int32_t* split_cnt, int32_t* sched, int64_t n_split,
                                        int64_t n_part, float* part_ms, float* part_acc, float* agg, int64_t ldo,
                                        void* agg_split, int opf, void* stream) {
  WSI_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31) && n_items >= 0 && n_items < (1ll << 31), "hetero_attn_work_fwd: bad sizes");
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(k && v && q && node_inv_r && e_w && e_b && (agg || agg_split) && items, "hetero_attn_work_fwd: null pointer");
  WSI_CHECK_ARG((reinterpret_cast<uintptr_t>(agg_split) & 7) == 0, "hetero_attn_work_fwd: agg_split must be 8 B aligned");
  WSI_CHECK_ARG(H >= 1 && D >= 1 && D % H == 0 && vec_ok(D, H),
                "hetero_attn_work_fwd: needs the lane-grouped layout (D %% 128 == 0, D <= 1024, H a power of two <= 32), got D=%d H=%d", D, H);
  WSI_CHECK_ARG(kv_dtype >= 0 && kv_dtype <= 2 && q_dtype >= 0 && q_dtype <= 2,
                "hetero_attn_work_fwd: unknown K / V / Q storage type %d / %d", kv_dtype, q_dtype);
  WSI_CHECK_ARG(ldk % 4 == 0 && ldv % 4 == 0 && ldq % 4 == 0 && ldo % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(k) & 7) == 0 && (reinterpret_cast<uintptr_t>(v) & 7) == 0,
                "hetero_attn_work_fwd: row strides must be multiples of 4 elements, K / V 8 B aligned");
  WSI_CHECK_ARG(n_split == 0 || (split_row && split_ptr && part_rel && part_ms && part_acc && n_part > 0),
                "hetero_attn_work_fwd: split rows need the partial buffers");
  WSI_CHECK_ARG(!split_cnt || part_split, "hetero_attn_work_fwd: split_cnt needs part_split");
  AttnArgs a{};
  a.K = reinterpret_cast<const float*>(k); a.ldk = ldk; a.V = reinterpret_cast<const float*>(v); a.ldv = ldv;
  a.Q = reinterpret_cast<const float*>(q); a.ldq = ldq;
  a.kv_dtype = kv_dtype; a.q_dtype = q_dtype;
  a.e_src = e_src; a.e_sim = e_sim; a.e_rel = e_rel; a.inv_r = node_inv_r;
  a.e_w = e_w; a.e_b = e_b; a.n_items = (int)n_items; a.D = D; a.H = H; a.dk = D / H;
  a.inv_sqrt_dk = 1.0f / sqrtf((float)(D / H));
  a.out = agg; a.ldo = ldo; a.attn = nullptr;
  a.n_src_rows = n_src_rows > 0 ? n_src_rows : n_rows;
  a.items = reinterpret_cast<const int4*>(items); a.part_ms = part_ms; a.part_acc = part_acc;
  WSI_CHECK_ARG(opf == WSI_OPF_BF16X3 || opf == WSI_OPF_F16 || opf == WSI_OPF_BF16, "hetero_attn_work_fwd: unknown operand format %d", opf);
  a.out_split = reinterpret_cast<__nv_bfloat16*>(agg_split);
  a.split_lo = opf == WSI_OPF_BF16X3 ? n_rows * D : (opf == WSI_OPF_F16 ? 0 : -1);
  if (n_split > 0 && split_cnt) {
    a.part_split = part_split; a.split_cnt = split_cnt;
    a.split_row = split_row; a.split_ptr = split_ptr; a.part_rel = part_rel;
  }
  a.sched = wsi_dev()->attn_static ? nullptr : sched;                       // development knob: static round-robin
  a.dbg = wsi_dev()->attn_debug;
  bool fused = false;
  int rc = launch<MODE_HEAT>(a, 1, wsi_stream(stream), &fused);
  if (rc != WSI_OK || fused) return rc;
  MergeArgs m{};
  m.split_row = split_row; m.split_ptr = split_ptr; m.part_rel = part_rel; m.part_ms = part_ms; m.part_acc = part_acc;
  m.inv_r = node_inv_r; m.n_split = (int)n_split; m.D = D; m.out = agg; m.ldo = ldo;
  m.out_split = a.out_split; m.split_lo = a.split_lo;
  return launch_merge(m, wsi_stream(stream));
}

extern "C" int wsi_hetero_attn_seg_fwd(const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_dtype, const void* qseg,
                                       int q_dtype, int64_t ldq, const int32_t* seg_ptr, const int32_t* seg_rel,
                                       const int32_t* e_src, const float* rel_pri, int64_t n_segs, int D, int H,
                                       int head_perm, const int32_t* items, int64_t n_src_rows, float* out, int64_t ldo,
                                       void* out_op, int opf, void* stream) {
  WSI_CHECK_ARG(kv_dtype >= 0 && kv_dtype <= 2, "hetero_attn_seg_fwd: unknown K / V storage type %d", kv_dtype);
  WSI_CHECK_ARG(n_segs >= 0 && n_segs < (1ll << 31), "hetero_attn_seg_fwd: bad n_segs");
  if (n_segs == 0) return WSI_OK;
  WSI_CHECK_ARG(k && v && qseg && (seg_ptr || items) && seg_rel && e_src && rel_pri && (out || out_op),
                "hetero_attn_seg_fwd: null pointer");
  WSI_CHECK_ARG(H >= 1 && D >= 1 && D % H == 0, "hetero_attn_seg_fwd: D=%d is not a multiple of H=%d", D, H);
  WSI_CHECK_ARG(head_perm || (!items && !out_op && out && q_dtype == 0),
                "hetero_attn_seg_fwd: the work list, 16-bit queries and the operand-form output need the lane-grouped layout (head_perm)");
  WSI_CHECK_ARG(q_dtype >= 0 && q_dtype <= 2, "hetero_attn_seg_fwd: unknown query storage type %d", q_dtype);
  WSI_CHECK_ARG(!out_op || (opf >= 0 && opf <= 2 && D % 8 == 0), "hetero_attn_seg_fwd: bad operand format %d", opf);
  AttnArgs a{};
  a.K = reinterpret_cast<const float*>(k); a.ldk = ldk; a.V = reinterpret_cast<const float*>(v); a.ldv = ldv; a.Q = reinterpret_cast<const float*>(qseg); a.ldq = ldq;
  a.kv_dtype = kv_dtype; a.q_dtype = q_dtype;
  a.n_src_rows = n_src_rows > 0 ? n_src_rows : 0;
  a.rowptr = seg_ptr; a.e_src = e_src; a.seg_rel = seg_rel; a.rel_pri = rel_pri;
  a.items = reinterpret_cast<const int4*>(items);
  a.n_items = (int)n_segs; a.D = D; a.H = H; a.dk = D / H;
  a.inv_sqrt_dk = 1.0f / sqrtf((float)(D / H));
  a.out = out; a.ldo = ldo;
  a.out_split = reinterpret_cast<__nv_bfloat16*>(out_op);
  a.split_lo = opf == WSI_OPF_BF16X3 ? n_segs * (int64_t)D : (opf == WSI_OPF_F16 ? 0 : -1);
  return launch<MODE_HGT_SEG>(a, head_perm, wsi_stream(stream));
}
