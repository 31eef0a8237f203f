// Whole-forward driver (wsi_heat_forward): the HEATNet2 / HEATNet4 inference chain as one host call.
// Replaces the Python-level sequencing of models/HEATNet4.py:195-247 (reference) / wsi_hgnn_b200/models/heat.py (ours):
// it only SEQUENCES the C-ABI kernels of this library on the caller's stream and carves their buffers out of one
// caller-owned workspace; no arithmetic lives here.
#include "common.cuh"

namespace {
inline int64_t al(int64_t v) { return (v + 1023) & ~(int64_t)1023; }

struct Carve {
  uintptr_t p, end;
  void* take(int64_t bytes) {
    p = (p + 1023) & ~(uintptr_t)1023;
    void* r = reinterpret_cast<void*>(p);
    p += bytes;
    return r;
  }
};
}  // namespace

extern "C" int64_t wsi_heat_forward_workspace_bytes(int64_t n_rows, int F, int D, int64_t n_part, int T, int B) {
  const int64_t N = n_rows;
  int64_t b = 2048;
  b += al(2 * N * F * 2);                 // feat [hi; lo]
  b += 2 * al(2 * N * D * 2);             // x [hi; lo], agg [hi; lo]
  b += al(N * 3 * D * 4);                 // K|V|Q
  b += 2 * al(N * D * 4);                 // x ping-pong
  b += al(n_part * 64 * 4) + al(n_part * D * 4);
  b += al(wsi_segment_pool_affine_workspace_bytes(N, (int64_t)T * B, D));
  return b;
}

extern "C" int wsi_heat_forward(const void* feat, int64_t ldf, int feat_is_op, const wsi_heat_graph* g,
                                const wsi_heat_params* p, float* x_out, int64_t ldx, float* logits, int64_t ldl,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  WSI_CHECK_ARG(g && p && feat && logits, "heat_forward: null pointer");
  const int64_t N = g->n_rows;
  const int T = g->T, B = g->B, F = p->F, D = p->D, H = p->H, L = p->L;
  WSI_CHECK_ARG(N > 0 && T >= 1 && T <= WSI_MAX_TYPES && B >= 1 && L >= 0 && g->type_ptr_host && g->type_ptr_host[T] == N,
                "heat_forward: bad graph (N=%lld T=%d B=%d)", (long long)N, T, B);
  WSI_CHECK_ARG(wsi_typed_linear_tc_ok(N, F, D) && (L == 0 || (wsi_typed_linear_tc_ok(N, D, 3 * D) && wsi_typed_linear_tc_ok(N, D, D))),
                "heat_forward: shapes (N=%lld F=%d D=%d) do not fit the tcgen05 chain", (long long)N, F, D);
  WSI_CHECK_ARG(p->n_out >= 1 && p->n_out <= 8 && p->M, "heat_forward: n_out=%d must be in [1, 8]", p->n_out);
  WSI_CHECK_ARG(!x_out || ldx >= D, "heat_forward: x_out row stride smaller than D");
  WSI_CHECK_ARG(ldf >= F && ldl >= p->n_out, "heat_forward: feat / logits row stride smaller than the row");
  const int opf = p->opf;
  WSI_CHECK_ARG(opf == WSI_OPF_BF16X3 || opf == WSI_OPF_F16 || opf == WSI_OPF_BF16, "heat_forward: unknown operand format %d", opf);
  WSI_CHECK_ARG(!feat_is_op || (ldf == F && (reinterpret_cast<uintptr_t>(feat) & 127) == 0),
                "heat_forward: operand-form features must be dense (ldf == F) and 128 B aligned");
  const int64_t m = opf == WSI_OPF_BF16X3 ? 2 : 1;       // 16-bit matrices per operand
  WSI_CHECK_ARG(p->w_in_split && g->seg_ptr && g->node_inv_r && g->e_src && g->e_sim && g->e_rel && g->items &&
                    (L == 0 || (p->w_kvq_split && p->b_kvq && p->w_a_split && p->b_a && p->skip && p->e_w && p->e_b)),
                "heat_forward: null pointer in the graph / parameter structs");
  for (int l = 0; l < L; ++l)
    WSI_CHECK_ARG(p->w_kvq_split[l] && p->w_a_split[l] && p->skip[l] && p->e_w[l] && p->e_b[l],
                  "heat_forward: null per-layer pointer (layer %d)", l);
  const int64_t need = wsi_heat_forward_workspace_bytes(N, F, D, g->n_part, T, B);
  WSI_CHECK_ARG(workspace && workspace_bytes >= need, "heat_forward: workspace of %lld bytes needed", (long long)need);

  Carve cv{reinterpret_cast<uintptr_t>(workspace), reinterpret_cast<uintptr_t>(workspace) + (uintptr_t)workspace_bytes};
  const void* feat_s = feat;
  if (!feat_is_op) feat_s = cv.take(m * N * F * 2);
  void* xs = cv.take(m * N * D * 2);
  void* aggs = cv.take(m * N * D * 2);
  float* kvq = static_cast<float*>(cv.take(N * 3 * D * 4));
  float* xa = static_cast<float*>(cv.take(N * D * 4));
  float* xb = static_cast<float*>(cv.take(N * D * 4));
  float* part_ms = static_cast<float*>(cv.take(g->n_part * 64 * 4));
  float* part_acc = static_cast<float*>(cv.take(g->n_part * D * 4));
  const int64_t pool_bytes = wsi_segment_pool_affine_workspace_bytes(N, (int64_t)T * B, D);
  void* pool_ws = cv.take(pool_bytes);
  const int32_t* tp = g->type_ptr_host;

  int rc = WSI_OK;
  if (!feat_is_op) rc = wsi_to_operand(static_cast<const float*>(feat), ldf, N, F, opf, const_cast<void*>(feat_s), stream);
  if (rc) return rc;
  // x = adapt_ws[type](feat)                                                   models/HEATNet4.py:198-206
  float* x = (L == 0 && x_out) ? x_out : xa;
  int64_t ld = (L == 0 && x_out) ? ldx : D;
  rc = wsi_typed_linear_op(feat_s, p->w_in_split, p->b_in, F, D, tp, T, WSI_ACT_NONE, nullptr, nullptr, 0, nullptr, 0,
                           nullptr, nullptr, x, ld, L > 0 ? xs : nullptr, opf, stream);
  if (rc) return rc;
  for (int l = 0; l < L; ++l) {                                                // models/HEATNet4.py:213-214
    rc = wsi_typed_linear_op(xs, p->w_kvq_split[l], p->b_kvq[l], D, 3 * D, tp, T, WSI_ACT_NONE, nullptr, nullptr, 0,
                             nullptr, 0, nullptr, nullptr, kvq, 3 * D, nullptr, opf, stream);   // :91-102, once per type
    if (rc) return rc;
    rc = wsi_hetero_attn_work_fwd(kvq, 3 * D, kvq + D, 3 * D, 0, kvq + 2 * D, 0, 3 * D, g->e_src, g->e_sim, g->e_rel,
                                  g->node_inv_r, p->e_w[l], p->e_b[l], N, N, D, H, g->items, g->n_items, g->split_row,
                                  g->split_ptr, g->part_rel, g->part_split, g->split_cnt, g->sched, g->n_split, g->n_part,
                                  part_ms, part_acc, nullptr, D, aggs, opf, stream);               // :103-119
    if (rc) return rc;
    const bool last = l + 1 == L;
    float* y = (last && x_out) ? x_out : (x == xa ? xb : xa);
    const int64_t ldy = (last && x_out) ? ldx : D;
    rc = wsi_typed_linear_op(aggs, p->w_a_split[l], p->b_a[l], D, D, tp, T, WSI_ACT_NONE, p->skip[l], x, ld, nullptr, 0,
                             g->node_inv_r, nullptr, y, ldy, last ? nullptr : xs, opf, stream);    // :121-136
    if (rc) return rc;
    x = y;
    ld = ldy;
  }
  // typed readout + linears_prediction (+ head_2 -> head_1 -> head, collapsed)   models/HEATNet4.py:216-245
  return wsi_segment_pool_affine_fwd(x, ld, g->seg_ptr, T, B, N, D, p->pool_op, p->M, p->c, p->b_total, p->seg_scale,
                                     p->n_out, 0, logits, ldl, pool_ws, pool_bytes, stream);
}

// ------------------------------------------------------------------------------------------------ blob -> logits
namespace {
struct SlideLayout {
  int64_t rowptr, e_src, e_sim, e_rel, stats, plan_ws, chunk_base, split_idx, hist, items, split_row, split_ptr, part_rel,
      part_split, zeroed, fwd, total;
};
SlideLayout slide_layout(int64_t N, int64_t E, int F, int D, int T, int chunk, int64_t max_part) {
  SlideLayout L{};
  int64_t o = 1024;
  auto take = [&](int64_t bytes) { int64_t r = o; o += al(bytes); return r; };
  L.rowptr = take((N + 1) * 4); L.e_src = take(E * 4); L.e_sim = take(E * 4); L.e_rel = take(E); L.stats = take(16);
  L.plan_ws = take(wsi_plan_workspace_bytes(N, E));
  L.chunk_base = take((N + 1) * 4); L.split_idx = take((N + 1) * 4); L.hist = take(2 * (chunk + 1) * 4);
  L.items = take((max_part + N + 1) * 16); L.split_row = take((N + 1) * 4); L.split_ptr = take((N + 2) * 4);
  L.part_rel = take((max_part + 1) * 4); L.part_split = take((max_part + 1) * 4); L.zeroed = take((N + 64) * 4);
  L.fwd = take(wsi_heat_forward_workspace_bytes(N, F, D, max_part, T, 1));
  L.total = o;
  return L;
}
}  // namespace

extern "C" int64_t wsi_slide_forward_workspace_bytes(int64_t n_nodes, int64_t n_edges, int F, int D, int T, int64_t max_part) {
  return slide_layout(n_nodes, n_edges, F, D, T, 16, max_part).total + 1024;
}

// Phase 1 of a slide: CSR build + work-list counting on `plan_stream`, totals copied (asynchronously) into the caller's
// pinned totals_host.  No host synchronisation: the caller waits for plan_stream (an event recorded after this call)
// before phase 2 - in the streaming evaluator that wait is a whole slide old and never blocks.
extern "C" int wsi_slide_plan(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, int32_t* totals_host,
                              void* workspace, int64_t workspace_bytes, void* plan_stream) {
  WSI_CHECK_ARG(s && p && totals_host && workspace, "slide_plan: null pointer");
  const int64_t N = s->n_nodes, E = s->n_edges;
  WSI_CHECK_ARG(N > 0 && E > 0 && N < (1ll << 31) && E < (1ll << 31) && s->chunk == 16 && max_part >= 1,
                "slide_plan: needs a non-empty slide, chunk 16 (N=%lld E=%lld chunk=%d)", (long long)N, (long long)E, s->chunk);
  const SlideLayout L = slide_layout(N, E, p->F, p->D, s->T, s->chunk, max_part);
  WSI_CHECK_ARG(workspace_bytes >= L.total + 1024, "slide_plan: workspace of %lld bytes needed", (long long)(L.total + 1024));
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  auto at = [&](int64_t off) { return base + off; };
  int32_t* rowptr = reinterpret_cast<int32_t*>(at(L.rowptr));
  int32_t* e_src = reinterpret_cast<int32_t*>(at(L.e_src));
  float* e_sim = reinterpret_cast<float*>(at(L.e_sim));
  uint8_t* e_rel = at(L.e_rel);
  int32_t* stats = reinterpret_cast<int32_t*>(at(L.stats));
  int32_t* chunk_base = reinterpret_cast<int32_t*>(at(L.chunk_base));
  int32_t* split_idx = reinterpret_cast<int32_t*>(at(L.split_idx));
  int32_t* hist = reinterpret_cast<int32_t*>(at(L.hist));
  cudaStream_t ps = wsi_stream(plan_stream);
  int rc = wsi_plan_build_csr(s->src, s->dst, s->sim, nullptr, s->rel_table, s->R, N, E, rowptr, e_src, e_sim, e_rel, nullptr,
                              stats, at(L.plan_ws), wsi_plan_workspace_bytes(N, E), plan_stream);
  if (rc) return rc;
  rc = wsi_plan_attn_work_count(rowptr, e_rel, N, s->chunk, chunk_base, split_idx, hist, at(L.plan_ws),
                                wsi_plan_workspace_bytes(N, 0), plan_stream);
  if (rc) return rc;
  WSI_CHECK_CUDA(cudaMemcpyAsync(totals_host, chunk_base + N, 4, cudaMemcpyDeviceToHost, ps));
  WSI_CHECK_CUDA(cudaMemcpyAsync(totals_host + 1, split_idx + N, 4, cudaMemcpyDeviceToHost, ps));
  WSI_CHECK_CUDA(cudaMemcpyAsync(totals_host + 2, stats, 8, cudaMemcpyDeviceToHost, ps));
  return WSI_OK;
}

// Phase 2: totals_host is valid - the caller has WAITED ON THE HOST for plan_stream to pass the copies of phase 1 - :
// work-list fill and forward on `stream`.
extern "C" int wsi_slide_run(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, const int32_t* totals_host,
                             float* logits, int64_t ldl, void* workspace, int64_t workspace_bytes, void* plan_stream,
                             void* stream) {
  WSI_CHECK_ARG(s && p && totals_host && logits && workspace, "slide_run: null pointer");
  const int64_t N = s->n_nodes, E = s->n_edges;
  WSI_CHECK_ARG(plan_stream != stream, "slide_run: plan_stream and stream must differ");
  const SlideLayout L = slide_layout(N, E, p->F, p->D, s->T, s->chunk, max_part);
  WSI_CHECK_ARG(workspace_bytes >= L.total + 1024, "slide_run: workspace of %lld bytes needed", (long long)(L.total + 1024));
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  auto at = [&](int64_t off) { return base + off; };
  int32_t* rowptr = reinterpret_cast<int32_t*>(at(L.rowptr));
  int32_t* e_src = reinterpret_cast<int32_t*>(at(L.e_src));
  float* e_sim = reinterpret_cast<float*>(at(L.e_sim));
  uint8_t* e_rel = at(L.e_rel);
  int32_t* chunk_base = reinterpret_cast<int32_t*>(at(L.chunk_base));
  int32_t* split_idx = reinterpret_cast<int32_t*>(at(L.split_idx));
  int32_t* hist = reinterpret_cast<int32_t*>(at(L.hist));
  cudaStream_t ps = wsi_stream(plan_stream), ms = wsi_stream(stream);
  const int64_t n_part = totals_host[0], n_split = totals_host[1];
  if (totals_host[3] != 0) { wsi_set_error("slide_forward: edge endpoint out of range"); return WSI_ERR_ARG; }
  WSI_CHECK_ARG(n_part >= 0 && n_split >= 0 && n_split <= N && n_part <= max_part,
                "slide_forward: %lld chunk partials exceed the workspace capacity %lld", (long long)n_part, (long long)max_part);
  int32_t* items = reinterpret_cast<int32_t*>(at(L.items));
  int32_t* zeroed = reinterpret_cast<int32_t*>(at(L.zeroed));
  // The fill runs on the MAIN stream: the caller has observed (on the host) that phase 1 finished, so nothing on
  // plan_stream needs to be waited for - and the next slide's phase 1, already queued on plan_stream, overlaps this
  // slide's forward instead of sitting in front of its work-list fill (measured: 0.48 -> 0.40 ms / slide streamed).
  int rc = wsi_plan_attn_work_fill(rowptr, e_rel, N, s->chunk, chunk_base, split_idx, n_part, n_split, hist, items,
                                   reinterpret_cast<int32_t*>(at(L.split_row)), reinterpret_cast<int32_t*>(at(L.split_ptr)),
                                   reinterpret_cast<int32_t*>(at(L.part_rel)), reinterpret_cast<int32_t*>(at(L.part_split)),
                                   stream);
  if (rc) return rc;
  WSI_CHECK_CUDA(cudaMemsetAsync(zeroed, 0, (size_t)(N + 64) * 4, ms));   // arrival counters of the fused merge, queue words
  (void)ps;

  wsi_heat_graph g{};
  g.n_rows = N; g.T = s->T; g.B = 1; g.type_ptr_host = s->type_ptr_host; g.seg_ptr = s->seg_ptr;
  g.e_src = e_src; g.e_sim = e_sim; g.e_rel = e_rel; g.node_inv_r = s->node_inv_r;
  g.items = items; g.n_items = n_part + N - n_split;
  g.split_row = reinterpret_cast<int32_t*>(at(L.split_row)); g.split_ptr = reinterpret_cast<int32_t*>(at(L.split_ptr));
  g.part_rel = reinterpret_cast<int32_t*>(at(L.part_rel)); g.part_split = reinterpret_cast<int32_t*>(at(L.part_split));
  g.split_cnt = zeroed; g.sched = zeroed + N + 62;
  g.n_split = n_split; g.n_part = n_part;
  return wsi_heat_forward(s->feat, s->ldf, s->feat_is_op, &g, p, nullptr, 0, logits, ldl, at(L.fwd),
                          wsi_heat_forward_workspace_bytes(N, p->F, p->D, max_part, s->T, 1), stream);
}

// Both phases in one call, with the one host synchronisation of plan_stream in between.
extern "C" int wsi_slide_forward(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, int32_t* totals_host,
                                 float* logits, int64_t ldl, void* workspace, int64_t workspace_bytes, void* plan_stream,
                                 void* stream) {
  WSI_CHECK_ARG(logits && plan_stream != stream, "slide_forward: plan_stream and stream must differ (the call synchronises plan_stream)");
  int rc = wsi_slide_plan(s, p, max_part, totals_host, workspace, workspace_bytes, plan_stream);
  if (rc) return rc;
  WSI_CHECK_CUDA(cudaStreamSynchronize(wsi_stream(plan_stream)));       // the one host read of the planner
  return wsi_slide_run(s, p, max_part, totals_host, logits, ldl, workspace, workspace_bytes, plan_stream, stream);
}
