// Edge builder kernels (K6 / K7): exact k-NN in feature space and per-edge Pearson correlation.
// Replaces Hnsw.fit / Hnsw.query (nmslib HNSW, approximate) and the per-edge scipy.pearsonr Python loop of the
// reference: construct_graph/graph_constructor.py:55-81, 262-282.
//
// k-NN = (1) dot products of a chunk of query rows with ALL rows by the typed-linear GEMM (tensor cores when the
// shape allows), (2) a streaming per-row selection of the 32 best candidates by the expanded form
// ||b||^2 - 2 a.b (warp-resident sorted list, one slot per lane), (3) an exact fp64 direct-form re-rank of those
// candidates ordered by (distance, index), (4) a MARGIN CHECK that makes the result exact by construction, not by
// luck: every node outside the shortlist has an approximate score >= the worst kept one, so its true score is
// >= worst_kept - eps (eps bounds the fp32 / split-product error of the expanded form); unless that still exceeds the
// exact topn-th score, the row is recomputed by exact fp64 brute force over all nodes (near-duplicate patches, features
// with a large common offset).  The emitted neighbour lists therefore always equal the brute-force answer.
#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CAND = 32;          // candidates kept per query row (>= topn + slack)

__global__ void __launch_bounds__(256)
row_sqnorm_kernel(const float* __restrict__ feat, int64_t n, int F, float* __restrict__ out, unsigned* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  float mx = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * 8) {
    const float* p = feat + row * F;
    float s = 0.f;
    for (int c = lane; c < F; c += 32) { float v = __ldg(p + c); s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) out[row] = s;
    mx = fmaxf(mx, s);
  }
  // largest ||b||^2 of the matrix (the error bound of the margin check): non-negative floats order like their bit patterns
  if (lane == 0 && mx > 0.f) atomicMax(max_bits, __float_as_uint(mx));
}

__device__ __forceinline__ bool pair_less(float d1, int i1, float d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

// one warp per query row: scan dot[q, 0..n) and keep the CAND smallest (||b||^2 - 2 a.b, index) pairs sorted across
// the lanes; then re-rank them by the exact fp64 distance and emit the first `topn`.
__global__ void __launch_bounds__(256)
knn_select_kernel(const float* __restrict__ feat, const float* __restrict__ sqn, const float* __restrict__ dot,
                  int64_t ld_dot, int64_t n, int F, int topn, int64_t q0, int n_q, int32_t* __restrict__ nbr,
                  float* __restrict__ nbr_dist, const unsigned* __restrict__ max_bits) {
  const int lane = threadIdx.x & 31;
  const float max_sq = __uint_as_float(__ldg(max_bits));
  for (int qi = blockIdx.x * 8 + (threadIdx.x >> 5); qi < n_q; qi += gridDim.x * 8) {
    const float* drow = dot + (int64_t)qi * ld_dot;
    float best_d = INFINITY;       // lane i holds the i-th smallest pair so far
    int best_i = 0x7fffffff;
    float thresh = INFINITY;       // = pair held by lane CAND-1
    int thresh_i = 0x7fffffff;
    // insertion of the candidates flagged in `hits` (lane l offers (d, base_idx + l * idx_stride)) into the sorted list
    auto offer = [&](unsigned hits, float d, int base_idx, int idx_stride) {
      while (hits) {
        const int src_lane = __ffs(hits) - 1;
        hits &= hits - 1;
        const float xd = __shfl_sync(FULL, d, src_lane);
        const int xi = base_idx + src_lane * idx_stride;
        if (!pair_less(xd, xi, thresh, thresh_i)) continue;      // threshold moved since the ballot
        const unsigned smaller = __ballot_sync(FULL, pair_less(best_d, best_i, xd, xi));
        const int pos = __popc(smaller);                          // list is sorted: `smaller` is a prefix mask
        const float up_d = __shfl_up_sync(FULL, best_d, 1);
        const int up_i = __shfl_up_sync(FULL, best_i, 1);
        if (lane > pos) { best_d = up_d; best_i = up_i; }
        else if (lane == pos) { best_d = xd; best_i = xi; }
        thresh = __shfl_sync(FULL, best_d, CAND - 1);
        thresh_i = __shfl_sync(FULL, best_i, CAND - 1);
      }
    };
    // Scan, 256 candidates per step: two 16-byte loads of the dot row and of the squared norms per lane, issued ONE STEP
    // AHEAD of their use, one
    // ballot that asks "does ANY of them beat the threshold" - after the first few hundred candidates almost every step
    // ends there (expected insertions per row ~ CAND ln(n / CAND)), so the scan runs at the rate the rows stream in.
    // (One candidate per lane and step, as in round 1, left one dependent L2 / DRAM round trip per 32 candidates:
    //  2.0 ms per 2560 x 100k chunk, 12x its bytes.)  The set of the CAND smallest (score, index) pairs does not depend
    // on the order in which candidates are offered, so the result is unchanged.
    const bool vec_ok = (ld_dot & 3) == 0 && (reinterpret_cast<uintptr_t>(drow) & 15) == 0 && (reinterpret_cast<uintptr_t>(sqn) & 15) == 0;
    const int64_t n_vec = vec_ok ? (n & ~(int64_t)255) : 0;
    float4 dv[2], sv[2], dn[2], sn[2];                            // current step / next step (loads issued one step ahead)
    if (n_vec > 0) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        dn[u] = __ldg(reinterpret_cast<const float4*>(drow + u * 128 + lane * 4));
        sn[u] = __ldg(reinterpret_cast<const float4*>(sqn + u * 128 + lane * 4));
      }
    }
    for (int64_t base = 0; base < n_vec; base += 256) {
#pragma unroll
      for (int u = 0; u < 2; ++u) { dv[u] = dn[u]; sv[u] = sn[u]; }
      if (base + 256 < n_vec) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          dn[u] = __ldg(reinterpret_cast<const float4*>(drow + base + 256 + u * 128 + lane * 4));
          sn[u] = __ldg(reinterpret_cast<const float4*>(sqn + base + 256 + u * 128 + lane * 4));
        }
      }
      float d[2][4];
      bool any = false;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        d[u][0] = fmaf(-2.f, dv[u].x, sv[u].x); d[u][1] = fmaf(-2.f, dv[u].y, sv[u].y);
        d[u][2] = fmaf(-2.f, dv[u].z, sv[u].z); d[u][3] = fmaf(-2.f, dv[u].w, sv[u].w);
#pragma unroll
        for (int k = 0; k < 4; ++k) any |= d[u][k] <= thresh;
      }
      if (!__ballot_sync(FULL, any)) continue;
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c0 = (int)base + u * 128 + k;                 // lane l holds candidate c0 + 4 l
          const unsigned hits = __ballot_sync(FULL, pair_less(d[u][k], c0 + 4 * lane, thresh, thresh_i));
          offer(hits, d[u][k], c0, 4);
        }
    }
    for (int64_t base = n_vec; base < n; base += 32) {
      const int64_t c = base + lane;
      float d = INFINITY;
      if (c < n) d = fmaf(-2.f, __ldg(drow + c), __ldg(sqn + c));
      const unsigned hits = __ballot_sync(FULL, c < n && pair_less(d, (int)c, thresh, thresh_i));
      offer(hits, d, (int)base, 1);
    }
    // exact re-rank: candidate j (held by lane j) gets its fp64 direct-form distance, computed by the whole warp
    const float* a = feat + (q0 + qi) * F;
    double my_d = INFINITY;
    const int n_cand = n < CAND ? (int)n : CAND;
    for (int j = 0; j < n_cand; ++j) {
      const int cj = __shfl_sync(FULL, best_i, j);
      const float* b = feat + (int64_t)cj * F;
      double s = 0.0;
      for (int c = lane; c < F; c += 32) { double df = (double)__ldg(a + c) - (double)__ldg(b + c); s = fma(df, df, s); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      if (lane == j) my_d = s;
    }
    int rank = 0;
    for (int j = 0; j < n_cand; ++j) {
      const double dj = __shfl_sync(FULL, my_d, j);
      const int ij = __shfl_sync(FULL, best_i, j);
      if (dj < my_d || (dj == my_d && ij < best_i)) ++rank;
    }
    // margin check (only when nodes were left out of the shortlist)
    bool exact_ok = true;
    if (n > CAND) {
      double aa = 0.0;
      for (int c = lane; c < F; c += 32) { const double v = (double)__ldg(a + c); aa = fma(v, v, aa); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) aa += __shfl_xor_sync(FULL, aa, o);
      // exact topn-th squared distance among the shortlist, as a score ||b||^2 - 2 a.b = ||a - b||^2 - ||a||^2
      const unsigned who = __ballot_sync(FULL, lane < n_cand && rank == topn - 1);
      const double dn = __shfl_sync(FULL, my_d, who ? __ffs(who) - 1 : 0);
      const double eps = ldexp(aa + (double)max_sq, -13);       // >= |fp32 expanded form - true score| for every node
      exact_ok = who != 0 && (double)thresh - eps > dn - aa + eps;
    }
    if (exact_ok) {
      if (lane < n_cand && rank < topn) {
        nbr[(int64_t)qi * topn + rank] = best_i;
        if (nbr_dist) nbr_dist[(int64_t)qi * topn + rank] = (float)sqrt(my_d);
      }
      continue;
    }
    // fallback: exact fp64 brute force of this row over ALL nodes, (distance, index) order; lane i holds the i-th best
    double bd = INFINITY;
    int bi = 0x7fffffff;
    for (int64_t c = 0; c < n; ++c) {
      const float* b = feat + c * F;
      double sdist = 0.0;
      for (int k = lane; k < F; k += 32) { const double df = (double)__ldg(a + k) - (double)__ldg(b + k); sdist = fma(df, df, sdist); }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sdist += __shfl_xor_sync(FULL, sdist, o);
      const double last_d = __shfl_sync(FULL, bd, topn - 1);
      const int last_i = __shfl_sync(FULL, bi, topn - 1);
      if (!(sdist < last_d || (sdist == last_d && (int)c < last_i))) continue;      // warp-uniform
      const unsigned smaller = __ballot_sync(FULL, bd < sdist || (bd == sdist && bi < (int)c));
      const int pos = __popc(smaller);
      const double up_d = __shfl_up_sync(FULL, bd, 1);
      const int up_i = __shfl_up_sync(FULL, bi, 1);
      if (lane > pos) { bd = up_d; bi = up_i; }
      else if (lane == pos) { bd = sdist; bi = (int)c; }
    }
    if (lane < topn) {
      nbr[(int64_t)qi * topn + lane] = bi;
      if (nbr_dist) nbr_dist[(int64_t)qi * topn + lane] = (float)sqrt(bd);
    }
  }
}

// Pearson r of two feature rows per edge: one warp per edge, fp32 loads, fp64 two-pass accumulation.
__global__ void __launch_bounds__(256)
edge_pearson_kernel(const float* __restrict__ feat, int F, const int64_t* __restrict__ src,
                    const int64_t* __restrict__ dst, int64_t n_edges, float* __restrict__ sim,
                    uint8_t* __restrict__ etype) {
  const int lane = threadIdx.x & 31;
  for (int64_t e = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); e < n_edges; e += (int64_t)gridDim.x * 8) {
    const float* a = feat + src[e] * F;
    const float* b = feat + dst[e] * F;
    double sa = 0.0, sb = 0.0;
    for (int c = lane; c < F; c += 32) { sa += (double)__ldg(a + c); sb += (double)__ldg(b + c); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(FULL, sa, o); sb += __shfl_xor_sync(FULL, sb, o); }
    const double ma = sa / F, mb = sb / F;
    double ab = 0.0, aa = 0.0, bb = 0.0;
    for (int c = lane; c < F; c += 32) {
      const double x = (double)__ldg(a + c) - ma, y = (double)__ldg(b + c) - mb;
      ab = fma(x, y, ab); aa = fma(x, x, aa); bb = fma(y, y, bb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ab += __shfl_xor_sync(FULL, ab, o); aa += __shfl_xor_sync(FULL, aa, o); bb += __shfl_xor_sync(FULL, bb, o);
    }
    if (lane == 0) {
      double r = ab / (sqrt(aa) * sqrt(bb));
      // a constant feature row has no correlation: scipy.stats.pearsonr returns NaN (ConstantInputWarning) and the
      // reference stores it, typing the edge 'neg' because `NaN > 0` is False (graph_constructor.py:278-281) - keep NaN
      // (fmin / fmax would silently turn it into -1); finite values are clipped to [-1, 1] as scipy does
      if (r == r) r = fmin(1.0, fmax(-1.0, r));
      sim[e] = (float)r;
      if (etype) etype[e] = r > 0.0 ? 1 : 0;        // 'pos' = 1, 'neg' = 0 (graph_constructor.py:281,296)
    }
  }
}

int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }

int64_t query_chunk(int64_t n, int64_t n_q) {
  int64_t c = (1ll << 28) / (n > 0 ? n : 1);      // <= 1 GiB of fp32 dot products per chunk
  c = c / 128 * 128;
  if (c < 128) c = 128;
  return c < n_q ? c : n_q;
}

}  // namespace

extern "C" int64_t wsi_knn_workspace_bytes(int64_t n, int F, int topn, int64_t q_begin, int64_t q_end) {
  (void)topn;
  if (n <= 0 || q_end <= q_begin) return 0;
  const int64_t qc = query_chunk(n, q_end - q_begin);
  const int64_t n4 = (n + 3) & ~(int64_t)3;       // row pitch of the dot-product matrix (16 B rows for the tcgen05 epilogue)
  return align256(n * 4 + 16) + align256(qc * n4 * 4) + align256(wsi_typed_linear_workspace_bytes(qc, F, (int)n, 1, 0, WSI_OPF_BF16X3));
}

extern "C" int wsi_knn_topk(const float* feat, int64_t n, int F, int topn, int64_t q_begin, int64_t q_end,
                            int32_t* nbr, float* nbr_dist, void* workspace, int64_t workspace_bytes, void* stream) {
  WSI_CHECK_ARG(n >= 0 && n < (1ll << 31) && F >= 1, "knn_topk: bad n / F");
  WSI_CHECK_ARG(0 <= q_begin && q_begin <= q_end && q_end <= n, "knn_topk: bad query range");
  if (q_end == q_begin) return WSI_OK;
  WSI_CHECK_ARG(topn >= 1 && topn <= CAND - 8, "knn_topk: topn must be in [1, %d]", CAND - 8);
  // the reference fails the slide when HNSW returns fewer than `radius` hits (np.stack -> ValueError,
  // graph_constructor.py:268-272 / get_graph.py:293-294)
  WSI_CHECK_ARG(topn <= n, "knn_topk: fewer than topn=%d nodes (n=%lld)", topn, (long long)n);
  WSI_CHECK_ARG(feat && nbr && workspace, "knn_topk: null pointer");
  WSI_CHECK_ARG(workspace_bytes >= wsi_knn_workspace_bytes(n, F, topn, q_begin, q_end), "knn_topk: workspace too small");
  cudaStream_t st = wsi_stream(stream);
  const int64_t qc = query_chunk(n, q_end - q_begin);
  char* ws = (char*)workspace;
  float* sqn = (float*)ws;
  const int64_t n4 = (n + 3) & ~(int64_t)3;
  unsigned* max_bits = (unsigned*)(ws + n * 4);              // one word behind the squared norms
  float* dot = (float*)(ws + align256(n * 4 + 16));
  void* lin_ws = ws + align256(n * 4 + 16) + align256(qc * n4 * 4);
  const int64_t lin_ws_bytes = wsi_typed_linear_workspace_bytes(qc, F, (int)n, 1, 0, WSI_OPF_BF16X3);
  const int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  const int64_t grid_cap = (int64_t)sms * 16;
  int blocks = (int)((n + 7) / 8 < grid_cap ? (n + 7) / 8 : grid_cap);
  WSI_CHECK_CUDA(cudaMemsetAsync(max_bits, 0, 4, st));
  row_sqnorm_kernel<<<blocks, 256, 0, st>>>(feat, n, F, sqn, max_bits);
  WSI_CHECK_LAUNCH();
  // query chunks of equal size (a short last chunk would fall off the tensor-core path: 160 rows of a 100k-node slide
  // cost 2 ms on the SIMT GEMM); the candidate matrix (all rows) is converted to the operand form ONCE
  const int64_t total_q = q_end - q_begin;
  const int64_t n_chunks = (total_q + qc - 1) / qc;
  int64_t chunk = (total_q + n_chunks - 1) / n_chunks;
  chunk = (chunk + 7) / 8 * 8;
  if (chunk > qc) chunk = qc;
  const bool tc = wsi_typed_linear_tc_ok(chunk < total_q ? chunk : total_q, F, (int)n) && n % 4 == 0 && total_q >= 512;
  char* a_ws = nullptr; char* w_ws = nullptr;
  if (tc) {
    uintptr_t wsp = (reinterpret_cast<uintptr_t>(lin_ws) + 1023) & ~(uintptr_t)1023;
    a_ws = reinterpret_cast<char*>(wsp);
    w_ws = reinterpret_cast<char*>(wsp + align256(2 * qc * (int64_t)F * 2));
    int rc = wsi_to_operand(feat, F, n, F, WSI_OPF_BF16X3, w_ws, stream);
    if (rc != WSI_OK) return rc;
  }
  for (int64_t q0 = q_begin; q0 < q_end; q0 += chunk) {
    const int n_q = (int)((q_end - q0) < chunk ? (q_end - q0) : chunk);
    int32_t tp[2] = {0, n_q};
    int rc;
    if (tc && wsi_typed_linear_tc_ok(n_q, F, (int)n)) {
      rc = wsi_to_operand(feat + q0 * F, F, n_q, F, WSI_OPF_BF16X3, a_ws, stream);
      if (rc != WSI_OK) return rc;
      rc = wsi_typed_linear_op(a_ws, w_ws, nullptr, F, (int)n, tp, 1, WSI_ACT_NONE, nullptr, nullptr, 0, nullptr, 0, nullptr,
                               nullptr, dot, n4, nullptr, WSI_OPF_BF16X3, stream);
    } else {
      rc = wsi_typed_linear_f32(feat + q0 * F, F, feat, nullptr, F, (int)n, tp, 1, WSI_ACT_NONE, nullptr, nullptr, 0,
                                nullptr, 0, nullptr, nullptr, dot, n4, 0, WSI_OPF_BF16X3, lin_ws, lin_ws_bytes, stream);
    }
    if (rc != WSI_OK) return rc;
    int sb = (int)((n_q + 7) / 8 < grid_cap ? (n_q + 7) / 8 : grid_cap);
    knn_select_kernel<<<sb, 256, 0, st>>>(feat, sqn, dot, n4, n, F, topn, q0, n_q, nbr + (q0 - q_begin) * topn,
                                          nbr_dist ? nbr_dist + (q0 - q_begin) * topn : nullptr, max_bits);
    WSI_CHECK_LAUNCH();
  }
  return WSI_OK;
}

extern "C" int wsi_edge_pearson(const float* feat, int64_t n, int F, const int64_t* src, const int64_t* dst,
                                int64_t n_edges, float* sim, uint8_t* etype, void* stream) {
  WSI_CHECK_ARG(n_edges >= 0 && F >= 1 && n >= 0, "edge_pearson: bad sizes");
  if (n_edges == 0) return WSI_OK;
  WSI_CHECK_ARG(feat && src && dst && sim, "edge_pearson: null pointer");
  const int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  int64_t want = (n_edges + 7) / 8;
  int blocks = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
  edge_pearson_kernel<<<blocks, 256, 0, wsi_stream(stream)>>>(feat, F, src, dst, n_edges, sim, etype);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
