// Native streaming evaluator: the per-epoch evaluation loop of the reference (evaluator/eval_homo_graph.py:61-95: for
// every slide `gnn(g.to(device))` and a synchronous read of the prediction) as ONE host call over a list of flat slides
// in pinned host memory.  Three CUDA streams, three slides in flight:
//     copy stream : slide i+2   ONE host -> device copy of the blob (+ its small plan head)
//     plan stream : slide i+1   wsi_slide_plan (CSR build + work-list counting, totals -> pinned host)
//     main stream : slide i     wsi_slide_run (work-list fill + the whole forward), logits -> pinned host
// The host thread only issues work: per slide ~20 launches and 3 copies, no Python, no allocation; the one wait per slide
// is on the plan event of a slide whose plan was enqueued a whole iteration earlier.  The plan head (segment pointers,
// relation table, 1/R per row) is derived here from the slide's header counts - nothing is carried over between slides.
#include <vector>

#include "common.cuh"

namespace {
inline int64_t al256(int64_t v) { return (v + 255) & ~(int64_t)255; }

struct HeadLayout { int64_t n0, n1, n1p, total_ints; };
HeadLayout head_layout(int64_t N, int T, int R) {
  HeadLayout h;
  h.n0 = T + 1;
  h.n1 = h.n0 + 3 * (R + 1);
  h.n1p = (h.n1 + T + 3) / 4 * 4;
  h.total_ints = h.n1p + N;
  return h;
}
}  // namespace

extern "C" int64_t wsi_stream_slot_bytes(int64_t max_nbytes, int64_t max_nodes, int64_t max_edges, int F, int D, int T, int R,
                                         int n_out) {
  return al256(max_nbytes) + al256(head_layout(max_nodes, T, R).total_ints * 4) + al256((int64_t)n_out * 4) +
         al256(wsi_slide_forward_workspace_bytes(max_nodes, max_edges, F, D, T, max_edges)) + 1024;
}

extern "C" int64_t wsi_stream_host_slot_bytes(int64_t max_nodes, int T, int R) {
  return al256(head_layout(max_nodes, T, R).total_ints * 4) + 256;
}

extern "C" int wsi_stream_forward(const wsi_stream_slide* slides, int64_t n_slides, const wsi_heat_params* p,
                                  float* logits_host, int depth, void* dev_ws, int64_t dev_ws_bytes, void* host_ws,
                                  int64_t host_ws_bytes, void* stream) {
  WSI_CHECK_ARG(n_slides >= 0 && depth >= 3 && depth <= 16, "stream_forward: depth must be in [3, 16]");
  if (n_slides == 0) return WSI_OK;
  WSI_CHECK_ARG(slides && p && logits_host && dev_ws && host_ws, "stream_forward: null pointer");
  int64_t max_nbytes = 0, max_nodes = 0, max_edges = 0;
  int T = slides[0].T, R_max = 0;
  for (int64_t i = 0; i < n_slides; ++i) {
    const wsi_stream_slide& s = slides[i];
    WSI_CHECK_ARG(s.blob_host && s.n_nodes > 0 && s.n_edges > 0 && s.T == T && s.T >= 1 && s.T <= WSI_MAX_TYPES && s.R >= 1 &&
                      s.R <= 255 && s.F == p->F && s.nodes_per_type_host && s.edges_per_rel_host && s.rel_src_type_host &&
                      s.rel_dst_type_host,
                  "stream_forward: slide %lld is not a non-empty flat slide of this model's shape", (long long)i);
    max_nbytes = s.nbytes > max_nbytes ? s.nbytes : max_nbytes;
    max_nodes = s.n_nodes > max_nodes ? s.n_nodes : max_nodes;
    max_edges = s.n_edges > max_edges ? s.n_edges : max_edges;
    R_max = s.R > R_max ? s.R : R_max;
  }
  const int64_t slot_bytes = wsi_stream_slot_bytes(max_nbytes, max_nodes, max_edges, p->F, p->D, T, R_max, p->n_out);
  const int64_t hslot_bytes = wsi_stream_host_slot_bytes(max_nodes, T, R_max);
  WSI_CHECK_ARG(dev_ws_bytes >= depth * slot_bytes, "stream_forward: device workspace of %lld bytes needed",
                (long long)(depth * slot_bytes));
  WSI_CHECK_ARG(host_ws_bytes >= depth * hslot_bytes, "stream_forward: pinned host workspace of %lld bytes needed",
                (long long)(depth * hslot_bytes));

  cudaStream_t ms = wsi_stream(stream), cs = nullptr, ps = nullptr;
  WSI_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  WSI_CHECK_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> uploaded(depth), planned(depth), done(depth);
  std::vector<char> used(depth, 0);
  for (int k = 0; k < depth; ++k) {
    WSI_CHECK_CUDA(cudaEventCreateWithFlags(&uploaded[k], cudaEventDisableTiming));
    WSI_CHECK_CUDA(cudaEventCreateWithFlags(&planned[k], cudaEventDisableTiming));
    WSI_CHECK_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
  }
  // everything already queued on the caller's stream (weight packs, ...) precedes the first forward
  std::vector<wsi_slide_desc> desc(depth);
  std::vector<wsi_heat_params> prm(depth, *p);
  std::vector<std::vector<int32_t>> tptr(depth, std::vector<int32_t>(T + 1));
  uint8_t* dbase = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dev_ws) + 255) & ~(uintptr_t)255);
  uint8_t* hbase = reinterpret_cast<uint8_t*>(host_ws);
  const int64_t off_head = al256(max_nbytes), off_logits = off_head + al256(head_layout(max_nodes, T, R_max).total_ints * 4);
  const int64_t off_ws = off_logits + al256((int64_t)p->n_out * 4);
  const int64_t slide_ws_bytes = wsi_slide_forward_workspace_bytes(max_nodes, max_edges, p->F, p->D, T, max_edges);
  int rc = WSI_OK;
  const int dbg = wsi_dev()->stream_debug;

  auto upload = [&](int64_t i) -> int {
    const int k = (int)(i % depth);
    const wsi_stream_slide& s = slides[i];
    if (used[k]) WSI_CHECK_CUDA(cudaEventSynchronize(done[k]));          // slot's previous slide (depth slides ago) has finished
    used[k] = 1;
    // ---- plan head from the header counts (host, O(N) for the per-row 1/R)
    const HeadLayout h = head_layout(s.n_nodes, s.T, s.R);
    int32_t* head = reinterpret_cast<int32_t*>(hbase + k * hslot_bytes);
    int32_t* tp = tptr[k].data();
    tp[0] = 0;
    for (int t = 0; t < s.T; ++t) tp[t + 1] = tp[t] + s.nodes_per_type_host[t];
    if (tp[s.T] != s.n_nodes) { wsi_set_error("stream_forward: slide %lld: node counts do not add up", (long long)i); return WSI_ERR_ARG; }
    for (int t = 0; t <= s.T; ++t) head[t] = tp[t];
    int32_t* tab = head + h.n0;
    int64_t e = 0;
    for (int r = 0; r <= s.R; ++r) { tab[r] = (int32_t)e; if (r < s.R) e += s.edges_per_rel_host[r]; }
    if (e != s.n_edges) { wsi_set_error("stream_forward: slide %lld: edge counts do not add up", (long long)i); return WSI_ERR_ARG; }
    std::vector<int> rcount(s.T, 0);
    for (int r = 0; r < s.R; ++r) {
      const int st = s.rel_src_type_host[r], dt = s.rel_dst_type_host[r];
      if (st < 0 || st >= s.T || dt < 0 || dt >= s.T) { wsi_set_error("stream_forward: slide %lld: bad relation types", (long long)i); return WSI_ERR_ARG; }
      tab[(s.R + 1) + r] = tp[st];
      tab[2 * (s.R + 1) + r] = tp[dt];
      ++rcount[dt];
    }
    tab[(s.R + 1) + s.R] = 0;
    tab[2 * (s.R + 1) + s.R] = 0;
    float* flags = reinterpret_cast<float*>(head + h.n1);
    for (int t = 0; t < s.T; ++t) flags[t] = s.nodes_per_type_host[t] > 0 ? 1.f : 0.f;
    for (int64_t j = h.n1 + s.T; j < h.n1p; ++j) head[j] = 0;
    float* inv = reinterpret_cast<float*>(head + h.n1p);
    for (int t = 0; t < s.T; ++t) {
      const float v = rcount[t] > 0 ? 1.f / (float)rcount[t] : 0.f;
      for (int j = tp[t]; j < tp[t + 1]; ++j) inv[j] = v;
    }
    // ---- the two copies of the slide
    uint8_t* slot = dbase + k * slot_bytes;
    if (!(dbg & 1) || i < depth)
      WSI_CHECK_CUDA(cudaMemcpyAsync(slot, s.blob_host, (size_t)s.nbytes, cudaMemcpyHostToDevice, cs));
    WSI_CHECK_CUDA(cudaMemcpyAsync(slot + off_head, head, (size_t)h.total_ints * 4, cudaMemcpyHostToDevice, cs));
    WSI_CHECK_CUDA(cudaEventRecord(uploaded[k], cs));
    // ---- descriptor of the device-resident slide
    wsi_slide_desc& d = desc[k];
    d = wsi_slide_desc{};
    const int32_t* hd = reinterpret_cast<const int32_t*>(slot + off_head);
    d.feat = slot + s.off_feat; d.ldf = s.F; d.feat_is_op = s.feat_is_op;
    d.src = reinterpret_cast<const int64_t*>(slot + s.off_src); d.dst = reinterpret_cast<const int64_t*>(slot + s.off_dst);
    d.sim = reinterpret_cast<const float*>(slot + s.off_sim);
    d.seg_ptr = hd; d.rel_table = hd + h.n0; d.node_inv_r = reinterpret_cast<const float*>(hd + h.n1p);
    d.type_ptr_host = tp;
    d.n_nodes = s.n_nodes; d.n_edges = s.n_edges; d.T = s.T; d.R = s.R; d.chunk = 16;
    prm[k] = *p;
    prm[k].seg_scale = reinterpret_cast<const float*>(hd + h.n1);
    return WSI_OK;
  };
  auto totals_of = [&](int k) { return reinterpret_cast<int32_t*>(hbase + k * hslot_bytes + al256(head_layout(max_nodes, T, R_max).total_ints * 4)); };
  auto plan = [&](int64_t i) -> int {
    const int k = (int)(i % depth);
    WSI_CHECK_CUDA(cudaStreamWaitEvent(ps, uploaded[k], 0));
    if (dbg & 2) { WSI_CHECK_CUDA(cudaEventRecord(planned[k], ps)); return WSI_OK; }
    int r = wsi_slide_plan(&desc[k], &prm[k], slides[i].n_edges, totals_of(k), dbase + k * slot_bytes + off_ws, slide_ws_bytes, ps);
    if (r) return r;
    WSI_CHECK_CUDA(cudaEventRecord(planned[k], ps));
    return WSI_OK;
  };
  auto run = [&](int64_t i) -> int {
    const int k = (int)(i % depth);
    WSI_CHECK_CUDA(cudaEventSynchronize(planned[k]));                   // recorded one iteration ago: the totals are on the host
    float* lg = reinterpret_cast<float*>(dbase + k * slot_bytes + off_logits);
    if (dbg & 2) { WSI_CHECK_CUDA(cudaStreamWaitEvent(ms, planned[k], 0)); WSI_CHECK_CUDA(cudaEventRecord(done[k], ms)); return WSI_OK; }
    int r = wsi_slide_run(&desc[k], &prm[k], slides[i].n_edges, totals_of(k), lg, p->n_out, dbase + k * slot_bytes + off_ws,
                          slide_ws_bytes, ps, ms);
    if (r) return r;
    WSI_CHECK_CUDA(cudaMemcpyAsync(logits_host + i * p->n_out, lg, (size_t)p->n_out * 4, cudaMemcpyDeviceToHost, ms));
    WSI_CHECK_CUDA(cudaEventRecord(done[k], ms));
    return WSI_OK;
  };

  // software pipeline: upload(i + 2) | plan(i + 1) | run(i)
  if ((rc = upload(0)) == WSI_OK && n_slides > 1) rc = upload(1);
  if (rc == WSI_OK) rc = plan(0);
  for (int64_t i = 0; i < n_slides && rc == WSI_OK; ++i) {
    if (i + 2 < n_slides) rc = upload(i + 2);
    if (rc == WSI_OK && i + 1 < n_slides) rc = plan(i + 1);
    if (rc == WSI_OK) rc = run(i);
  }
  cudaError_t e1 = cudaStreamSynchronize(ms), e2 = cudaStreamSynchronize(ps), e3 = cudaStreamSynchronize(cs);
  for (int k = 0; k < depth; ++k) { cudaEventDestroy(uploaded[k]); cudaEventDestroy(planned[k]); cudaEventDestroy(done[k]); }
  cudaStreamDestroy(cs);
  cudaStreamDestroy(ps);
  if (rc != WSI_OK) return rc;
  WSI_CHECK_CUDA(e1);
  WSI_CHECK_CUDA(e2);
  WSI_CHECK_CUDA(e3);
  return WSI_OK;
}
