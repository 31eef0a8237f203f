// Graph plan builder (kernel K8): the device layout the message-passing kernels consume, built on the device.
//
//   wsi_plan_build_csr  : per-relation COO (local node ids, as a DGL heterograph stores them) -> one dst-major CSR over
//                         the type-packed node ids whose rows hold their in-edges grouped by relation, original
//                         edge order kept inside a (dst, relation) segment.  Replaces what dgl.to_heterogeneous and
//                         DGL's on-demand CSC conversion do for the reference
//                         (construct_graph/graph_constructor.py:285-297; the per-relation sub_graph views of
//                         models/HEATNet4.py:91-92).
//   wsi_plan_attn_work_*: the hub-balancing work list of wsi_hetero_attn_work_fwd.
//
// Integer / index work, HBM-latency bound and tiny next to the forward (E * ~40 B); what matters is that it is a
// handful of launches instead of ~40 framework ops on the end-to-end path.  Deterministic: a row is filled with
// atomics and then rank-sorted by the original edge position, so the layout does not depend on the atomic order.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---------------------------------------------------------------------------------------------- exclusive scan
// out[i] = sum_{j<i} in[j] for i in [0, n]; (out has n + 1 entries, out[n] = total).  Three launches.
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  __shared__ int block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
    if (lane == SCAN_THREADS / 32 - 1) block_total = wi;
  }
  __syncthreads();
  *total = block_total;
  const int r = warp_sums[warp] + incl - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const int* in, int64_t n, int* out, int* tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n ? in[base + i] : 0; s += v[i]; }
  int total;
  int off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = off; off += v[i]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums (n_tiles <= SCAN_TILE), total to out_total
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int* tile_sums, int n_tiles, int* out_total) {
  const int base = threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = base + i < n_tiles ? tile_sums[base + i] : 0; s += v[i]; }
  int total;
  int off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n_tiles) tile_sums[base + i] = off; off += v[i]; }
  if (threadIdx.x == 0) *out_total = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int* out, int64_t n, const int* tile_sums) {
  const int add = tile_sums[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) out[base + i] += add;
}

// scan of in[0..n) into out[0..n]; tile_sums: scratch of ceil(n / SCAN_TILE) ints
int exclusive_scan(const int* in, int64_t n, int* out, int* tile_sums, cudaStream_t stream) {
  const int64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (n_tiles > SCAN_TILE) { wsi_set_error("plan: more than %d nodes are not supported", SCAN_TILE * SCAN_TILE); return WSI_ERR_UNSUPPORTED; }
  if (n == 0) { WSI_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(int), stream)); return WSI_OK; }
  scan_tiles_kernel<<<(int)n_tiles, SCAN_THREADS, 0, stream>>>(in, n, out, tile_sums);
  WSI_CHECK_LAUNCH();
  scan_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(tile_sums, (int)n_tiles, out + n);
  WSI_CHECK_LAUNCH();
  if (n_tiles > 1) {
    scan_add_kernel<<<(int)n_tiles, SCAN_THREADS, 0, stream>>>(out, n, tile_sums);
    WSI_CHECK_LAUNCH();
  }
  return WSI_OK;
}

// ---------------------------------------------------------------------------------------------- CSR build
// rel_table int32 [3, R + 1]: edge range of relation r in the concatenated arrays | src type offset | dst type offset
__device__ __forceinline__ int rel_of_edge(const int* rel_ptr, int R, int64_t e) {
  int lo = 0, hi = R - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(rel_ptr + mid) <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) csr_count_kernel(const int64_t* src, const int64_t* dst, const int* rel_table, int R,
                                                        int64_t n_nodes, int64_t n_edges, int* count, int* stats) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int r = rel_of_edge(rel_table, R, e);
  const int64_t s = src[e] + __ldg(rel_table + (R + 1) + r), d = dst[e] + __ldg(rel_table + 2 * (R + 1) + r);
  if (s < 0 || s >= n_nodes || d < 0 || d >= n_nodes) { atomicOr(stats + 1, 1); return; }   // edge endpoint out of range
  atomicAdd(count + d, 1);
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int64_t* dst, const int* rel_table, int R, int64_t n_nodes,
                                                       int64_t n_edges, const int* rowptr, int* cursor, int* slot_edge) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int r = rel_of_edge(rel_table, R, e);
  const int64_t d = dst[e] + __ldg(rel_table + 2 * (R + 1) + r);
  if (d < 0 || d >= n_nodes) return;
  const int pos = atomicAdd(cursor + d, 1);
  slot_edge[__ldg(rowptr + d) + pos] = (int)e;
}

// one warp per row: rank-sort the row's edge positions (ascending = relation, then original order) and emit the payload
__global__ void __launch_bounds__(256) csr_emit_kernel(const int64_t* src, const float* sim, const double* sim64,
                                                       const int* rel_table, int R, int64_t n_nodes, const int* rowptr,
                                                       const int* slot_edge, int* e_src, float* e_sim, uint8_t* e_rel,
                                                       int* e_dst, int* stats) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_nodes) return;
  const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1), n = end - beg;
  if (n == 0) return;
  if (lane == 0) atomicMax(stats, n);
  for (int i = lane; i < n; i += 32) {
    const int e = __ldg(slot_edge + beg + i);
    int rank = 0;
    if (n <= 32) {
      const unsigned mask = n == 32 ? 0xffffffffu : ((1u << n) - 1u);             // exactly the lanes < n are here
      for (int j = 0; j < n; ++j) rank += __shfl_sync(mask, e, j) < e;
    } else {
      for (int j = 0; j < n; ++j) rank += __ldg(slot_edge + beg + j) < e;
    }
    const int r = rel_of_edge(rel_table, R, e);
    const int o = beg + rank;
    e_src[o] = (int)(src[e] + __ldg(rel_table + (R + 1) + r));
    e_sim[o] = sim ? sim[e] : (sim64 ? (float)sim64[e] : 0.f);
    e_rel[o] = (uint8_t)r;
    if (e_dst) e_dst[o] = (int)row;
  }
}

// ---------------------------------------------------------------------------------------------- attention work list
// chunks of a split row = sum over its (row, relation) segments of ceil(len / chunk); 0 for rows with <= chunk edges
// class_hist int32 [2 * (chunk + 1)]: [0, chunk] = number of whole-row items with chunk - c edges (largest first);
// the second half is the fill cursor of each class (zeroed here, used by work_fill_kernel)
__global__ void __launch_bounds__(256) work_count_kernel(const int* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                                                         int* row_chunks, int* row_split, int* class_hist) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_nodes) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  int chunks = 0;
  if (end - beg > chunk) {
    int seg_len = 0, rel = -1;
    for (int e = beg; e < end; ++e) {
      const int r = e_rel[e];
      if (r != rel) { chunks += (seg_len + chunk - 1) / chunk; seg_len = 0; rel = r; }
      ++seg_len;
    }
    chunks += (seg_len + chunk - 1) / chunk;
  }
  row_chunks[row] = chunks;
  row_split[row] = chunks > 0;
  if (chunks == 0) atomicAdd(class_hist + (chunk - (end - beg)), 1);
}

__global__ void __launch_bounds__(256) work_fill_kernel(const int* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                                                        const int* chunk_base, const int* split_idx, int n_part,
                                                        int n_split, int* class_hist, int4* items, int* split_row,
                                                        int* split_ptr, int* part_rel, int* part_split) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row == 0 && n_split >= 0) split_ptr[n_split] = n_part;
  if (row >= n_nodes) return;
  const int beg = rowptr[row], end = rowptr[row + 1];
  const int si = split_idx[row];
  if (end - beg <= chunk) {
    // whole row: after the chunk items, LARGEST FIRST (the kernel deals items to its warps round-robin, so a
    // size-sorted list balances them).  The order inside a size class is arbitrary (atomic cursor): it only
    // decides which warp runs the item, never a result.
    const int c = chunk - (end - beg);
    int base = n_part;
    for (int i = 0; i < c; ++i) base += class_hist[i];
    const int pos = base + atomicAdd(class_hist + (chunk + 1) + c, 1);
    items[pos] = make_int4((int)row, beg, end, -1);
    return;
  }
  int slot = chunk_base[row];
  split_row[si] = (int)row;
  split_ptr[si] = slot;
  int e = beg;
  while (e < end) {
    const int rel = e_rel[e];
    int seg_end = e + 1;
    while (seg_end < end && e_rel[seg_end] == rel) ++seg_end;
    for (int c = e; c < seg_end; c += chunk) {
      items[slot] = make_int4((int)row, c, min(c + chunk, seg_end), slot);
      part_rel[slot] = rel;
      if (part_split) part_split[slot] = si;
      ++slot;
    }
    e = seg_end;
  }
}

inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }

}  // namespace

// workspace layout: count/cursor int32 [N + 1] | slot_edge int32 [E] | tile sums
extern "C" int64_t wsi_plan_workspace_bytes(int64_t n_nodes, int64_t n_edges) {
  return align256((n_nodes + 1) * 4) * 2 + align256(n_edges * 4) + align256(((n_nodes + SCAN_TILE) / SCAN_TILE + 1) * 4) + 256;
}

extern "C" int wsi_plan_build_csr(const int64_t* src, const int64_t* dst, const float* sim, const double* sim64,
                                  const int32_t* rel_table, int R, int64_t n_nodes, int64_t n_edges, int32_t* rowptr,
                                  int32_t* e_src, float* e_sim, uint8_t* e_rel, int32_t* e_dst, int32_t* stats,
                                  void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = wsi_stream(stream_);
  WSI_CHECK_ARG(n_nodes >= 0 && n_edges >= 0 && n_nodes < (1ll << 31) && n_edges < (1ll << 31), "plan_build_csr: bad sizes");
  WSI_CHECK_ARG(rowptr && stats, "plan_build_csr: null pointer");
  WSI_CHECK_ARG(R >= 0 && R <= 255, "plan_build_csr: at most 255 relations (got %d)", R);
  WSI_CHECK_ARG(workspace_bytes >= wsi_plan_workspace_bytes(n_nodes, n_edges) && (workspace || workspace_bytes == 0),
                "plan_build_csr: workspace too small");
  WSI_CHECK_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(int), stream));
  if (n_edges == 0 || R == 0) {
    WSI_CHECK_CUDA(cudaMemsetAsync(rowptr, 0, (n_nodes + 1) * sizeof(int), stream));
    return WSI_OK;
  }
  WSI_CHECK_ARG(src && dst && rel_table && e_src && e_sim && e_rel, "plan_build_csr: null pointer");
  WSI_CHECK_ARG(!(sim && sim64), "plan_build_csr: give the edge attribute as fp32 or fp64, not both");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int* count = reinterpret_cast<int*>(ws);
  int* cursor = reinterpret_cast<int*>(ws + align256((n_nodes + 1) * 4));
  int* slot_edge = reinterpret_cast<int*>(ws + 2 * align256((n_nodes + 1) * 4));
  int* tile_sums = reinterpret_cast<int*>(ws + 2 * align256((n_nodes + 1) * 4) + align256(n_edges * 4));
  WSI_CHECK_CUDA(cudaMemsetAsync(ws, 0, 2 * align256((n_nodes + 1) * 4), stream));
  const int eb = (int)((n_edges + 255) / 256);
  csr_count_kernel<<<eb, 256, 0, stream>>>(src, dst, rel_table, R, n_nodes, n_edges, count, stats);
  WSI_CHECK_LAUNCH();
  int rc = exclusive_scan(count, n_nodes, rowptr, tile_sums, stream);
  if (rc != WSI_OK) return rc;
  csr_fill_kernel<<<eb, 256, 0, stream>>>(dst, rel_table, R, n_nodes, n_edges, rowptr, cursor, slot_edge);
  WSI_CHECK_LAUNCH();
  csr_emit_kernel<<<(int)((n_nodes + 7) / 8), 256, 0, stream>>>(src, sim, sim64, rel_table, R, n_nodes, rowptr, slot_edge,
                                                                e_src, e_sim, e_rel, e_dst, stats);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

// Phase 1: per-row chunk counts and their scans.  chunk_base / split_idx int32 [N + 1] (exclusive scans; the last
// entries are n_part and n_split - read them on the host to size the buffers of phase 2).
extern "C" int wsi_plan_attn_work_count(const int32_t* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                                        int32_t* chunk_base, int32_t* split_idx, int32_t* class_hist, void* workspace,
                                        int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = wsi_stream(stream_);
  WSI_CHECK_ARG(n_nodes >= 0 && n_nodes < (1ll << 31) && chunk >= 1 && chunk <= 255, "plan_attn_work_count: bad sizes (1 <= chunk <= 255)");
  WSI_CHECK_ARG(rowptr && chunk_base && split_idx && class_hist, "plan_attn_work_count: null pointer");
  WSI_CHECK_CUDA(cudaMemsetAsync(class_hist, 0, 2 * (chunk + 1) * sizeof(int), stream));
  WSI_CHECK_ARG(workspace_bytes >= wsi_plan_workspace_bytes(n_nodes, 0) && workspace, "plan_attn_work_count: workspace too small");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  int* row_chunks = reinterpret_cast<int*>(ws);
  int* row_split = reinterpret_cast<int*>(ws + align256((n_nodes + 1) * 4));
  int* tile_sums = reinterpret_cast<int*>(ws + 2 * align256((n_nodes + 1) * 4));
  if (n_nodes > 0) {
    work_count_kernel<<<(int)((n_nodes + 255) / 256), 256, 0, stream>>>(rowptr, e_rel, n_nodes, chunk, row_chunks, row_split, class_hist);
    WSI_CHECK_LAUNCH();
  }
  int rc = exclusive_scan(row_chunks, n_nodes, chunk_base, tile_sums, stream);
  if (rc != WSI_OK) return rc;
  return exclusive_scan(row_split, n_nodes, split_idx, tile_sums, stream);
}

// Phase 2: items int32 [n_part + (N - n_split), 4], split_row [n_split], split_ptr [n_split + 1], part_rel [n_part]
extern "C" int wsi_plan_attn_work_fill(const int32_t* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                                       const int32_t* chunk_base, const int32_t* split_idx, int64_t n_part,
                                       int64_t n_split, int32_t* class_hist, int32_t* items, int32_t* split_row,
                                       int32_t* split_ptr, int32_t* part_rel, int32_t* part_split, void* stream_) {
  cudaStream_t stream = wsi_stream(stream_);
  WSI_CHECK_ARG(n_nodes >= 0 && n_nodes < (1ll << 31) && chunk >= 1 && n_part >= 0 && n_split >= 0, "plan_attn_work_fill: bad sizes");
  if (n_nodes == 0) return WSI_OK;
  WSI_CHECK_ARG(rowptr && chunk_base && split_idx && items && split_ptr && class_hist, "plan_attn_work_fill: null pointer");
  WSI_CHECK_ARG(n_split == 0 || (split_row && part_rel), "plan_attn_work_fill: null pointer");
  work_fill_kernel<<<(int)((n_nodes + 255) / 256), 256, 0, stream>>>(rowptr, e_rel, n_nodes, chunk, chunk_base, split_idx,
                                                                    (int)n_part, (int)n_split, class_hist,
                                                                    reinterpret_cast<int4*>(items), split_row, split_ptr,
                                                                    part_rel, part_split);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
