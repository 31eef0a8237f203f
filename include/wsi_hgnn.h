/* wsi_hgnn.h - C ABI of libwsi_hgnn.so, the sm_100a CUDA library behind wsi_hgnn_b200.
 *
 * The reference (HKU-MedAI/WSI-HGNN) has no native code of its own: every native instruction of
 * its hot path runs inside DGL / cuBLAS / nmslib / scipy.  Each entry point below therefore cites
 * the reference *call site* it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - the caller (PyTorch caching allocator) owns every buffer, workspaces included; the library
 *    never allocates or frees device memory and keeps no reference to caller memory;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and nothing
 *    synchronises (graph-capturable);
 *  - return 0 on success, <0 on error (WSI_ERR_*); the message is in wsi_last_error()
 *    (thread-local).  No exception crosses this boundary.
 *  - node rows are packed TYPE-MAJOR: packed id = type_ptr[t] + local id.
 */
#ifndef WSI_HGNN_H_
#define WSI_HGNN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSI_ABI_VERSION 22

#define WSI_ERR_ARG (-1)
#define WSI_ERR_CUDA (-2)
#define WSI_ERR_UNSUPPORTED (-3)

/* activation codes of the typed-linear epilogue */
#define WSI_ACT_NONE 0
#define WSI_ACT_GELU 1 /* exact erf form == torch F.gelu default (models/HGT.py:180) */

/* readout ops: pooling/avg_pooling.py:15-17, sum_pooling.py:14-16, max_pooling.py:15-17 */
#define WSI_POOL_SUM 0
#define WSI_POOL_MEAN 1
#define WSI_POOL_MAX 2

/* Operand formats of the tensor-core typed linear (kernel K1) and of the buffers that feed it.
 *   WSI_OPF_BF16X3  bf16 [2 * rows, K]: rows [0, rows) = hi = bf16(x), rows [rows, 2 rows) = lo = bf16(x - hi); the
 *                   product is hi*hi + hi*lo + lo*hi (3 MMAs, ~2^-17 relative): "exact" mode, used for gradients
 *   WSI_OPF_F16     fp16 [rows, K], round to nearest, clamped to +-65504: ONE MMA per k-slice; 11-bit significand
 *                   (TF32's).  Default of the fp32 models: the whole forward stays 3.6x inside the 1e-3 parity bar on
 *                   the 16 reference-generated goldens and config 2 (profiles/r2_precision_study.json)
 *   WSI_OPF_BF16    bf16 [rows, K]: one MMA; the bf16-storage configuration (BASELINE config 3)
 */
#define WSI_OPF_BF16X3 0
#define WSI_OPF_F16 1
#define WSI_OPF_BF16 2

/* attention scoring modes */
#define WSI_SCORE_HEAT 0 /* score = <q,k> * (w*sim+b) / sqrt(d_k)        models/HEATNet4.py:103,111 */
#define WSI_SCORE_HGT 1  /* score = <q',k> * relation_pri[r,h] / sqrt(d_k) models/HGT.py:100         */

int wsi_abi_version(void);
const char* wsi_last_error(void);
/* number of SMs of the current device (148 on B200); <0 on error */
int wsi_num_sms(void);
/* make `device` the calling thread's current CUDA device for this library (its CUDA runtime is linked statically) */
int wsi_set_device(int device);
/* number of kernels this library has launched so far in this process (bench.py: gpu_launches) */
int64_t wsi_launch_count(void);
/* DEVELOPMENT hook, not product API: sets one of the process-wide kernel debug / variant knobs the library otherwise
 * reads once from the environment at load time (tools/sweep_dev.py); unknown key -> WSI_ERR_ARG. */
int wsi_dev_set(const char* key, int value);

/* ---------------------------------------------------------------------------------------------
 * Typed linear:  Y[rows of type t] = epilogue( X[rows of type t] . W[t]^T )        (kernel K1)
 * replaces the per-node-type nn.Linear calls  models/HEATNet4.py:100-102,134,202,219,243-245,
 * models/HEATNet2.py:75-77,109,166,188, models/HGT.py:82-84,121,180,194.
 *   x [N, ldx] fp32, w [T, n_out, K] fp32, bias [T, n_out] or NULL, type_ptr_host int32 [T+1].
 * epilogue, in this order (each part optional):
 *   v = acc + bias; v = act(v); v *= drop_mask[row, n];
 *   if skip: a = sigmoid(skip[t]);  v = row_gate[row]!=0 ? v*a + res[row,n]*(1-a) : res[row,n]
 *            (the sigma(skip) mix of models/HEATNet4.py:122-136 incl. its KeyError passthrough)
 *   v *= row_scale[row]
 * y [N, ldy] fp32.  `impl`: 0 = auto (tcgen05 tensor-core path when the shape is tile aligned,
 * else the fp32 SIMT path), 1 = force SIMT, 2 = force tcgen05 (error if the shape does not fit).
 * The tcgen05 path converts x and w to the operand format `opf` (WSI_OPF_*) in a pre-pass, accumulates in fp32 in
 * TMEM, and needs `workspace` of wsi_typed_linear_workspace_bytes() bytes.
 */
int64_t wsi_typed_linear_workspace_bytes(int64_t n_rows, int K, int n_out, int T, int impl, int opf);
int wsi_typed_linear_f32(const float* x, int64_t ldx, const float* w, const float* bias, int K, int n_out,
                         const int32_t* type_ptr_host, int T, int act, const float* skip, const float* res,
                         int64_t ldres, const float* drop_mask, int64_t ldmask, const float* row_gate,
                         const float* row_scale, float* y, int64_t ldy, int impl, int opf, void* workspace,
                         int64_t workspace_bytes, void* stream);

/* Operands kept in operand form for chains of tensor-core GEMMs (no per-call conversion pass):
 * wsi_to_operand: fp32 [rows, K] (row stride ld_src) -> dst in format `opf` (16-bit [2 * rows, K] for WSI_OPF_BF16X3,
 *   [rows, K] otherwise).  K % 8 == 0.  Weights [T, n_out, K] are converted as rows = T * n_out.
 * wsi_typed_linear_op: the typed linear of wsi_typed_linear_f32 on the tcgen05 path with x and w given in operand
 *   form (both 128 B aligned); y_op != NULL additionally emits the epilogue result in operand form ([2N, n_out] or
 *   [N, n_out], n_out % 8 == 0) for the next GEMM; y may be NULL then.
 *   Returns WSI_ERR_UNSUPPORTED when wsi_typed_linear_tc_ok(N, K, n_out) == 0. */
int wsi_typed_linear_tc_ok(int64_t n_rows, int K, int n_out);
int wsi_to_operand(const float* src, int64_t ld_src, int64_t rows, int K, int opf, void* dst, void* stream);
/* dst row i = operand form of src row row_idx[i] (int32 [rows]): gather + conversion in one pass - the (dst, relation)
 * segments of HGT pick up their dst node's query (models/HGT.py:88-92 moved to the dst side). */
int wsi_gather_to_operand(const float* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, int opf, void* dst,
                          void* stream);
/* src fp32 [N, K] -> WSI_OPF_BF16X3 operand form dst [2N, K] AND colsum [T, K] = per-type column sums, in one pass: the
 * backward of a per-type nn.Linear needs dY as a GEMM operand (data / weight gradient) and its column sums as the bias
 * gradient (trainer/train_gnn.py:68-71).  Deterministic (per-tile partials summed in a fixed order).
 * workspace: wsi_to_operand_colsum_workspace_bytes(K, type_ptr_host, T) bytes. */
int64_t wsi_to_operand_colsum_workspace_bytes(int K, const int32_t* type_ptr_host, int T);
int wsi_to_operand_colsum(const float* src, int64_t ld_src, int K, const int32_t* type_ptr_host, int T, void* dst,
                          float* colsum, void* workspace, int64_t workspace_bytes, void* stream);
/* The same gather on a matrix that already is in a single-plane operand form (WSI_OPF_F16 / WSI_OPF_BF16, [n, K] 16-bit,
 * row stride ld_src elements - a column slice of a wider matrix is fine): dst (dense) row i = src row row_idx[i]. */
int wsi_gather_rows16(const void* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, void* dst, void* stream);
int wsi_typed_linear_op(const void* x_op, const void* w_op, const float* bias, int K, int n_out,
                        const int32_t* type_ptr_host, int T, int act, const float* skip, const float* res,
                        int64_t ldres, const float* drop_mask, int64_t ldmask, const float* row_gate,
                        const float* row_scale, float* y, int64_t ldy, void* y_op, int opf, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Heterogeneous edge attention forward, ONE launch for all relations of a layer        (kernel K2)
 * replaces, per relation, apply_edges(v_dot_u) + score + edge_softmax + u_mul_e/sum and the
 * cross-relation mean of multi_update_all:
 *   models/HEATNet4.py:103-119 == models/HEATNet2.py:78-94  (WSI_SCORE_HEAT).
 * Graph layout (GraphPlan): CSR over packed dst rows, in-edges of a row grouped by relation slot.
 *   rowptr int32 [N+1], e_src int32 [E] (packed src row), e_sim fp32 [E], e_rel uint8 [E]
 *   node_inv_r fp32 [N] : 1/R_t of the row's dst type (0 => no incoming relation => row written as 0)
 * k/v rows are gathered by src (row strides ldk/ldv floats), q by dst.  agg [N, ldo]:
 *   agg[v,h,:] = inv_r[v] * sum_{segments (v,r)} softmax_{e in seg}(score_e,h) . V[src e,h,:]
 * `head_perm` != 0: the D columns of k,q,v,agg are in the "lane-grouped" physical order produced by
 * wsi_head_perm() (requires D % 128 == 0, H a power of two <= 32); 0: natural order, any D, H.
 * HEAT mode reads the e_linear scalars from device memory (e_w, e_b: 1 float each).
 * If attn_out != NULL it receives the per-(edge, head) normalised attention a[e,h] ([E, H], for backward).
 */
int wsi_hetero_attn_fwd(const float* k, int64_t ldk, const float* v, int64_t ldv, const float* q, int64_t ldq,
                        const int32_t* rowptr, const int32_t* e_src, const float* e_sim, const uint8_t* e_rel,
                        const float* node_inv_r, const float* e_w, const float* e_b, int64_t n_rows, int D,
                        int H, int head_perm, float* agg, int64_t ldo, float* attn_out, void* stream);

/* Same computation driven by a WORK LIST that balances the heavy-tailed in-degree of k-NN graphs (hubs):
 *   items int32 [n_items, 4] = (row, e_beg, e_end, slot).  slot < 0: [e_beg, e_end) is the whole row, written
 *   to agg[row] (zero in-degree rows need an item too).  slot >= 0: a chunk of ONE (row, relation) segment whose
 *   online-softmax partial goes to part_ms [n_part, 64] (per-lane max | sum) and part_acc [n_part, D].
 *   The split rows are finished by a merge launch: split_row int32 [n_split], split_ptr int32 [n_split + 1]
 *   (partial slots of each split row, in edge order), part_rel int32 [n_part] (relation slot of each partial).
 *   With split_cnt int32 [n_split] (ZERO before the first launch; the kernel leaves it zero) and part_split int32
 *   [n_part] the merge is fused: the warp finishing a row's last chunk merges that row; else a second launch does.
 *   sched int32 [2] (optional, ZERO before the first launch, left zero): device-side work queue - warps pull items in
 *   list order (largest first = LPT) instead of a static round-robin.
 * Requires the lane-grouped column order (head_perm layout of wsi_head_perm).  Built by GraphPlan.attn_work().
 *   kv_dtype: storage type of k / v - 0 fp32, 1 fp16, 2 bf16 (ldk / ldv in elements; same lane-grouped order): half the
 *   gathered bytes (and half the exchange of the node-sharded layer), fp32 scores / softmax / accumulation.
 *   q_dtype: storage type of q, same codes.  n_src_rows: rows of k / v (0 = n_rows; HGT runs this kernel over its
 *   (dst, relation) SEGMENTS as rows, whose count is not the node count): the K | V footprint picks the kernel.
 *   agg_op != NULL: the result is (also) written in operand format `opf` (WSI_OPF_*: 16-bit [2 * n_rows, D] hi rows then
 *   lo rows, or [n_rows, D]), the A operand of wsi_typed_linear_op; agg may then be NULL. */
int wsi_hetero_attn_work_fwd(const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_dtype, const void* q, int q_dtype,
                             int64_t ldq, const int32_t* e_src, const float* e_sim, const uint8_t* e_rel,
                             const float* node_inv_r, const float* e_w, const float* e_b, int64_t n_rows,
                             int64_t n_src_rows, int D, int H, const int32_t* items,
                             int64_t n_items, const int32_t* split_row, const int32_t* split_ptr,
                             const int32_t* part_rel, const int32_t* part_split, int32_t* split_cnt, int32_t* sched,
                             int64_t n_split, int64_t n_part, float* part_ms,
                             float* part_acc, float* agg, int64_t ldo, void* agg_op, int opf, void* stream);

/* Backward of wsi_hetero_attn_fwd (kernel K3; HEAT scoring, lane-grouped column order): what DGL's GSDDMM / EdgeSoftmax /
 * GSpMM backward compute under loss.backward() (trainer/train_gnn.py:68-71 through models/HEATNet4.py:103-119).
 *   d_agg [N, ldg] = gradient of agg.  dk, dv [N, ld*]: ACCUMULATED into (zero them first; rows are shared between
 *   destinations -> vector atomics); dq [N, lddq]: written; d_e [2] = (d e_linear.weight, d e_linear.bias): accumulated.
 *   Nothing is saved by the forward: the segment softmax is recomputed from k, v, q.
 *   row_order int32 [N] or NULL: processing order of the dst rows (largest in-degree first balances the k-NN hubs).
 *   Two-pass mode (t_ptr != NULL; no atomics, deterministic): the transposed (SOURCE-major) edge list of the same graph
 *   - t_ptr int32 [n_src + 1], t_eid int32 [E] (position of the edge in the dst-major arrays), t_dst int32 [E] (its dst
 *   row) - and coef_ws fp32 [E, 2, H].  The dst-major pass writes dq and the per-(edge, head) coefficients, a second
 *   kernel walks the source rows and WRITES dk / dv (all n_src rows; no zero fill needed). */
int wsi_hetero_attn_bwd(const float* k, int64_t ldk, const float* v, int64_t ldv, const float* q, int64_t ldq,
                        const int32_t* rowptr, const int32_t* e_src, const float* e_sim, const uint8_t* e_rel,
                        const float* node_inv_r, const float* e_w, const float* e_b, int64_t n_rows, int D, int H,
                        const float* d_agg, int64_t ldg, float* dk, int64_t lddk, float* dv, int64_t lddv, float* dq,
                        int64_t lddq, float* d_e, const int32_t* row_order, const int32_t* t_ptr, const int32_t* t_eid,
                        const int32_t* t_dst, int64_t n_src, float* coef_ws, void* stream);

/* Segment form used by HGT (WSI_SCORE_HGT): one work item per (dst,relation) segment.
 *   seg_ptr int32 [S+1] edge range of segment s (dst-major order), seg_rel int32 [S] MODEL relation id,
 *   qseg [S, ldq] = transformed query of the segment (Q[dst] . relation_att[r]^T per head,
 *   models/HGT.py:88-92 moved to the dst side), rel_pri [R_model, H] (models/HGT.py:59,100).
 *   out [S, ldo] = softmax-weighted sum of V[src] over the segment (before relation_msg).
 *   kv_dtype: storage type of k / v - 0 fp32, 1 fp16, 2 bf16 (ldk / ldv in elements).  The 16-bit forms are the
 *   bf16-storage configuration (BASELINE config 3): gathered bytes halve, scores / softmax / accumulation stay fp32;
 *   both kernels (head_perm 0 / 1) take them.  q_dtype: storage of qseg, same codes (16-bit: head_perm only) - the
 *   relation_att GEMM then writes only its 16-bit output.
 *   items int32 [n_segs, 4] or NULL (head_perm only): work list (row, e_beg, e_end, -1) - work item i reads qseg row
 *   `row`, seg_rel[row], the edges [e_beg, e_end) and writes out row `row`; this is how the relation-SORTED segment order
 *   of the tensor-core relation transforms runs over the dst-major edge arrays (seg_ptr may then be NULL).
 *   n_src_rows: rows of k / v that can be gathered (0 = unknown): the K | V footprint picks the register-path or the
 *   TMA bulk-copy ring kernel, as in wsi_hetero_attn_work_fwd.
 *   out_op (head_perm only) or NULL: operand-form copy of out (wsi_to_operand layout for `opf`, [n_segs, D]) - the A
 *   operand of the relation_msg GEMM; out may be NULL when only out_op is wanted.
 */
int wsi_hetero_attn_seg_fwd(const void* k, int64_t ldk, const void* v, int64_t ldv, int kv_dtype, const void* qseg,
                            int q_dtype, int64_t ldq, const int32_t* seg_ptr, const int32_t* seg_rel, const int32_t* e_src,
                            const float* rel_pri, int64_t n_segs, int D, int H, int head_perm, const int32_t* items,
                            int64_t n_src_rows, float* out, int64_t ldo, void* out_op, int opf, void* stream);

/* Physical column order used when head_perm != 0: perm_host[p] = logical column stored at physical
 * position p (int32 [D]).  Returns <0 if (D,H) has no lane-grouped layout. */
int wsi_head_perm(int D, int H, int32_t* perm_host);

/* ---------------------------------------------------------------------------------------------
 * HGT relation transforms (kernel K5)  models/HGT.py:88-93: per (relation, head) d_k x d_k maps, applied
 * to (dst, relation) SEGMENTS instead of to every node x relation.  Segments are processed grouped by
 * relation: grouped position i in [rel_ptr_host[r], rel_ptr_host[r+1]) belongs to model relation r.
 *   y[y_row_idx[i], h, :] = W[r,h] . x[x_row_idx[i], h, :]      (w_kn == 0:  y_n = sum_k W[n,k] x_k)
 *   y[y_row_idx[i], h, :] = x[x_row_idx[i], h, :] . W[r,h]      (w_kn != 0:  y_n = sum_k x_k W[k,n])
 * x [*, ldx], w [R, H, d_k, d_k], y [*, ldy]; x_row_idx / y_row_idx int32 [S] (NULL = identity).
 */
int wsi_rel_transform(const float* x, int64_t ldx, const int32_t* x_row_idx, const int32_t* y_row_idx,
                      const float* w, const int32_t* rel_ptr_host, int R, int H, int d_k, int w_kn, float* y,
                      int64_t ldy, void* stream);

/* Sum of the segment messages of each dst row, times 1/R_t:  the stack->mean of
 * multi_update_all(..., cross_reducer='mean')  models/HGT.py:105-106.
 *   row_seg_ptr int32 [N+1] segment range of row v; msg [S, ldm] of storage type msg_dtype (0 fp32, 1 fp16, 2 bf16: the
 *   16-bit output of the relation_msg GEMM); agg [N, ldo] (or NULL).
 *   seg_pos int32 [S] or NULL: msg row of segment s (the relation-sorted order of the tensor-core transforms).
 *   agg_op or NULL: operand-form copy of agg (wsi_to_operand layout for `opf`), the A operand of the a_linear GEMM. */
int wsi_segment_combine(const void* msg, int msg_dtype, int64_t ldm, const int32_t* row_seg_ptr, const int32_t* seg_pos,
                        const float* node_inv_r, int64_t n_rows, int D, float* agg, int64_t ldo, void* agg_op, int opf,
                        void* stream);

/* Typed LayerNorm  models/HGT.py:123-124 (nn.LayerNorm(out_dim) per node type, eps 1e-5), in place allowed.
 *   gamma/beta [T, D]; rows of type t = [type_ptr_host[t], type_ptr_host[t+1]).
 *   row_gate [N] or NULL: rows with gate == 0 are copied through un-normalised - the reference `continue`s before the
 *   norm for a node type without incoming relation (models/HGT.py:118-120); per ROW because in a pack()ed batch that is
 *   a per-(type, graph) property (pass node_inv_r).
 *   y_op or NULL (D % 128 == 0, D <= 1024): operand-form copy of the result (wsi_to_operand layout for `opf`), the A
 *   operand of the next layer's GEMMs - written in the same pass; y may then be NULL. */
int wsi_typed_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, const float* row_gate,
                        const int32_t* type_ptr_host, int T, int D, float eps, float* y, int64_t ldy, void* y_op, int opf,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * Graph plan builder (kernel K8): what dgl.to_heterogeneous + DGL's on-demand CSC conversion do for the reference
 * (construct_graph/graph_constructor.py:285-297; the per-relation sub_graph views of models/HEATNet4.py:91-92).
 * Input: the per-relation COO arrays concatenated in relation order, LOCAL node ids (as a heterograph stores them):
 *   src, dst int64 [E]; the edge attribute as sim fp32 [E] or sim64 fp64 [E] (old pickles, graph_constructor.py:292)
 *   or neither (0); rel_table (DEVICE) int32 [3, R + 1]: row 0 = edge range of relation r, row 1 = packed-id offset
 *   of its src type, row 2 = of its dst type.
 * Output: rowptr int32 [N + 1], e_src int32 [E], e_sim fp32 [E], e_rel uint8 [E], optional e_dst int32 [E]: edges
 *   sorted by (packed dst, relation, original position).  stats (DEVICE) int32 [4]: [0] max in-degree, [1] != 0 if an
 *   edge endpoint was out of range (such edges are dropped; the caller raises).  Deterministic. */
int64_t wsi_plan_workspace_bytes(int64_t n_nodes, int64_t n_edges);
int wsi_plan_build_csr(const int64_t* src, const int64_t* dst, const float* sim, const double* sim64,
                       const int32_t* rel_table, int R, int64_t n_nodes, int64_t n_edges, int32_t* rowptr,
                       int32_t* e_src, float* e_sim, uint8_t* e_rel, int32_t* e_dst, int32_t* stats, void* workspace,
                       int64_t workspace_bytes, void* stream);
/* Work list of wsi_hetero_attn_work_fwd in two phases (the host reads the two totals in between to size the arrays):
 *   count: chunk_base, split_idx int32 [N + 1] = exclusive scans of (chunks of row, row is split); last entries =
 *          n_part, n_split; class_hist int32 [2 * (chunk + 1)] scratch shared by the two phases (chunk <= 255).
 *          workspace: wsi_plan_workspace_bytes(N, 0).
 *   fill : items [n_part + N - n_split, 4] = the chunk items (edge order), then the whole-row items sorted by edge
 *          count, largest first (order inside a size class arbitrary: it only affects scheduling);
 *          split_row [n_split], split_ptr [n_split + 1], part_rel [n_part], part_split [n_part] (optional: split
 *          index of each partial, for the fused merge). */
int wsi_plan_attn_work_count(const int32_t* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                             int32_t* chunk_base, int32_t* split_idx, int32_t* class_hist, void* workspace,
                             int64_t workspace_bytes, void* stream);
int wsi_plan_attn_work_fill(const int32_t* rowptr, const uint8_t* e_rel, int64_t n_nodes, int chunk,
                            const int32_t* chunk_base, const int32_t* split_idx, int64_t n_part, int64_t n_split,
                            int32_t* class_hist, int32_t* items, int32_t* split_row, int32_t* split_ptr,
                            int32_t* part_rel, int32_t* part_split, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Typed readout (kernel K4): dgl.readout.{sum,mean,max}_nodes(graph, 'h', ntype=)  pooling/*.py
 *   x [N, ldx]; seg_ptr int32 [n_seg+1] row ranges of the (type, graph) segments; out [n_seg, ldo];
 *   empty segment -> 0.  workspace: wsi_segment_pool_workspace_bytes().
 */
int64_t wsi_segment_pool_workspace_bytes(int64_t n_rows, int64_t n_seg, int D);
int wsi_segment_pool_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t n_seg, int64_t n_rows,
                         int D, int op, float* out, int64_t ldo, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* Typed readout fused with the (narrow, affine) prediction that follows it:
 *   out[b, o] (+)= b_total[o] + sum_t seg_scale[t*B + b] * ( <M[t, o, :], pool(x rows of segment t*B + b)> + c[t, o] )
 * covers models/HEATNet2.py:181-194 (M = linears_prediction weights), the per-layer readout of models/HGT.py:189-199
 * (accumulate != 0) and, for inference, models/HEATNet4.py:216-245 whose linears_prediction -> cat -> head_2 ->
 * head_1 -> head chain has no nonlinearity and is collapsed on the host into one [out, D] map per node type.
 *   seg_ptr int32 [T*B + 1] type-major segments; M [T, n_out, D]; c [T, n_out] or NULL; b_total [n_out] or NULL;
 *   seg_scale [T*B] or NULL (0 => the reference's zero block for an empty node type); 1 <= n_out <= 8. */
int64_t wsi_segment_pool_affine_workspace_bytes(int64_t n_rows, int64_t n_seg, int D);
int wsi_segment_pool_affine_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int T, int B, int64_t n_rows, int D,
                                int op, const float* M, const float* c, const float* b_total, const float* seg_scale,
                                int n_out, int accumulate, float* out, int64_t ldo, void* workspace,
                                int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge builder (kernels K6/K7): construct_graph/graph_constructor.py:256-303.
 * wsi_knn_topk: exact k-NN in feature space, L2, self included, ordered by (distance, index);
 *   replaces Hnsw.fit/query (graph_constructor.py:55-81,262-273).  feat [N, F] fp32 row-major.
 *   For query rows [q_begin, q_end): nbr int32 [(q_end-q_begin), topn] = the topn nearest nodes
 *   (rank 0 normally the node itself; the caller drops it, graph_constructor.py:270).
 * wsi_edge_pearson: sim[e] = pearson(feat[src e], feat[dst e]) (graph_constructor.py:276-282),
 *   etype[e] = sim > 0.
 */
int64_t wsi_knn_workspace_bytes(int64_t n, int F, int topn, int64_t q_begin, int64_t q_end);
int wsi_knn_topk(const float* feat, int64_t n, int F, int topn, int64_t q_begin, int64_t q_end, int32_t* nbr,
                 float* nbr_dist, void* workspace, int64_t workspace_bytes, void* stream);
int wsi_edge_pearson(const float* feat, int64_t n, int F, const int64_t* src, const int64_t* dst, int64_t n_edges,
                     float* sim, uint8_t* etype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-forward driver: HEATNet2.forward / HEATNet4.forward (models/HEATNet4.py:195-247, models/HEATNet2.py:159-196)
 * in inference, as ONE host call that enqueues the kernel chain on `stream`
 *   split(feat) -> adapt_ws GEMM -> L x [ K|V|Q GEMM -> edge attention (work list) -> a_linear GEMM + skip mix ]
 *   -> fused typed readout + collapsed affine heads
 * (the same launches the Python modules issue one by one; here the host cost per slide is a few microseconds per
 * kernel, which is what bounds the streamed end-to-end path).  Preconditions = those of the tcgen05 chain:
 * wsi_typed_linear_tc_ok(n_rows, F, D), (n_rows, D, 3D), (n_rows, D, D); wsi_head_perm(D, H) exists; n_out <= 8.
 * All weights are in the forms the kernels consume: operand-format stacks (`opf`, wsi_to_operand) in the GRAPH's
 * node-type order, K|V|Q fused along the output dimension, K/V/Q rows and a_linear columns in head_perm order.
 */
typedef struct wsi_heat_graph {
  int64_t n_rows;                 /* N packed nodes */
  int32_t T, B;                   /* node types of the graph, graphs in the batch */
  const int32_t* type_ptr_host;   /* [T+1] HOST */
  const int32_t* seg_ptr;         /* [T*B+1] readout segments, type-major */
  const int32_t* e_src;           /* CSR arrays of wsi_plan_build_csr */
  const float* e_sim;
  const uint8_t* e_rel;
  const float* node_inv_r;
  const int32_t* items;           /* work list of wsi_plan_attn_work_fill */
  int64_t n_items;
  const int32_t* split_row;
  const int32_t* split_ptr;
  const int32_t* part_rel;
  const int32_t* part_split;
  int32_t* split_cnt;
  int32_t* sched;                 /* may be NULL */
  int64_t n_split, n_part;
} wsi_heat_graph;

typedef struct wsi_heat_params {
  int32_t F, D, H, L;             /* input / hidden width, heads, layers */
  int32_t opf;                    /* WSI_OPF_*: format of the w_*_op stacks and of the chain's internal operands */
  const void* w_in_split;         /* operand form of [T*D, F] */
  const float* b_in;              /* [T, D] */
  const void* const* w_kvq_split; /* HOST array [L] of the operand form of [T*3D, D] */
  const float* const* b_kvq;      /* HOST array [L] of [T, 3D] */
  const void* const* w_a_split;   /* HOST array [L] of the operand form of [T*D, D] */
  const float* const* b_a;        /* HOST array [L] of [T, D] */
  const float* const* skip;       /* HOST array [L] of [T] */
  const float* const* e_w;        /* HOST array [L] of device scalars (e_linear.weight) */
  const float* const* e_b;        /* HOST array [L] of device scalars (e_linear.bias) */
  int32_t pool_op, n_out;         /* WSI_POOL_*, logits width (<= 8) */
  const float* M;                 /* [T, n_out, D] collapsed prediction maps (wsi_segment_pool_affine_fwd) */
  const float* c;                 /* [T, n_out] or NULL */
  const float* b_total;           /* [n_out] or NULL */
  const float* seg_scale;         /* [T*B] or NULL */
} wsi_heat_params;

int64_t wsi_heat_forward_workspace_bytes(int64_t n_rows, int F, int D, int64_t n_part, int T, int B);
/* feat [N, ldf] type-major packed: fp32 when feat_is_op == 0; already in operand form p->opf (dense, ldf == F, 128 B
 * aligned - a flat slide that stores its features in fp16) when feat_is_op != 0.  x_out [N, ldx] final node embeddings
 * (head_perm column order is NOT applied to x: embeddings are in natural order) or NULL; logits [B, ldl]. */
int wsi_heat_forward(const void* feat, int64_t ldf, int feat_is_op, const wsi_heat_graph* g, const wsi_heat_params* p,
                     float* x_out, int64_t ldx, float* logits, int64_t ldl, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * One slide, blob to logits, in ONE host call: the per-slide body of the streaming evaluator
 * (evaluator/eval_homo_graph.py:61-95 drives `gnn(g.to(device))` slide by slide).  For a flat slide already resident
 * on the device (features | src | dst | sim, wsi_hgnn_b200/slide_io.py) it enqueues
 *     plan_stream:  wsi_plan_build_csr -> wsi_plan_attn_work_count -> [host read of the 4 totals] -> wsi_plan_attn_work_fill
 *     stream     :  waits for the plan (event), then wsi_heat_forward
 * EXCEPTION to the conventions at the top of this header: it synchronises `plan_stream` once (the totals size the work
 * list) - `plan_stream` must therefore carry nothing but this slide's upload dependency and plan.
 * workspace: wsi_slide_forward_workspace_bytes(); it is still in use by `stream` when the call returns (the caller
 * recycles it only after the slide's forward has completed).  max_part = capacity for hub-row chunk partials
 * (n_edges always suffices); a slide that needs more fails with WSI_ERR_ARG.
 */
typedef struct wsi_slide_desc {
  const void* feat;               /* [N, ldf] packed features: fp32, or (feat_is_op != 0) operand form p->opf, ldf == F */
  int64_t ldf;
  int32_t feat_is_op;
  const int64_t* src;             /* [E] local ids, relation-major */
  const int64_t* dst;
  const float* sim;               /* [E] or NULL */
  const int32_t* rel_table;       /* DEVICE int32 [3, R + 1] (wsi_plan_build_csr) */
  const int32_t* seg_ptr;         /* DEVICE int32 [T + 1] readout segments of the one graph (= type_ptr) */
  const float* node_inv_r;        /* DEVICE [N] */
  const int32_t* type_ptr_host;   /* HOST int32 [T + 1] */
  int64_t n_nodes, n_edges;
  int32_t T, R, chunk;
} wsi_slide_desc;

int64_t wsi_slide_forward_workspace_bytes(int64_t n_nodes, int64_t n_edges, int F, int D, int T, int64_t max_part);
/* totals_host: PINNED host int32 [4] scratch (n_part, n_split, max in-degree, bad-edge flag on return) */
int wsi_slide_forward(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, int32_t* totals_host,
                      float* logits, int64_t ldl, void* workspace, int64_t workspace_bytes, void* plan_stream,
                      void* stream);
/* The same in two phases WITHOUT any host synchronisation inside the library, for a software-pipelined caller:
 *   wsi_slide_plan : CSR build + work-list counting on plan_stream; the 4 totals are copied asynchronously to totals_host
 *   wsi_slide_run  : after the caller has WAITED ON THE HOST until plan_stream passed those copies (an event it
 *                    recorded after wsi_slide_plan - in the streaming evaluator that event is a whole slide old):
 *                    work-list fill and forward on `stream`.  Same desc / params / workspace in both calls. */
int wsi_slide_plan(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, int32_t* totals_host,
                   void* workspace, int64_t workspace_bytes, void* plan_stream);
int wsi_slide_run(const wsi_slide_desc* s, const wsi_heat_params* p, int64_t max_part, const int32_t* totals_host,
                  float* logits, int64_t ldl, void* workspace, int64_t workspace_bytes, void* plan_stream, void* stream);

/* The whole streaming evaluator natively: a LIST of flat slides in pinned host memory -> their logits, in one host call.
 * Per slide: ONE host -> device copy of the blob (+ the small plan head the call derives from the header counts),
 * wsi_slide_plan on a plan stream, wsi_slide_run + logits D2H on `stream`; three slides in flight on three streams
 * (two of them created and destroyed by the call), `depth` device slots.  Synchronises `stream` before returning.
 *   dev_ws: depth * wsi_stream_slot_bytes(max blob bytes, max nodes, max edges, F, D, T, max R, n_out) device bytes
 *   host_ws: depth * wsi_stream_host_slot_bytes(max nodes, T, max R) PINNED host bytes;  logits_host: pinned [n, n_out]
 *   p->seg_scale is ignored (taken from every slide's own head). */
typedef struct wsi_stream_slide {
  const void* blob_host;              /* pinned: features | src | dst | sim (FlatSlide layout) */
  int64_t nbytes;
  int64_t off_feat, off_src, off_dst, off_sim;   /* byte offsets inside the blob */
  int64_t n_nodes, n_edges;
  int32_t T, R, F, feat_is_op;        /* feat_is_op != 0: features stored in the operand format p->opf (fp16) */
  const int32_t* nodes_per_type_host; /* [T] */
  const int32_t* edges_per_rel_host;  /* [R] relation-major edge counts */
  const int32_t* rel_src_type_host;   /* [R] node-type index of the relation's sources */
  const int32_t* rel_dst_type_host;   /* [R] ... destinations */
} wsi_stream_slide;
int64_t wsi_stream_slot_bytes(int64_t max_nbytes, int64_t max_nodes, int64_t max_edges, int F, int D, int T, int R, int n_out);
int64_t wsi_stream_host_slot_bytes(int64_t max_nodes, int T, int R);
int wsi_stream_forward(const wsi_stream_slide* slides, int64_t n_slides, const wsi_heat_params* p, float* logits_host,
                       int depth, void* dev_ws, int64_t dev_ws_bytes, void* host_ws, int64_t host_ws_bytes, void* stream);

/* Backward of the fused a_linear epilogue (sigma(skip) mix + dropout mask + KeyError passthrough of
 * models/HEATNet4.py:122-136; forward = the epilogue of wsi_typed_linear_op) in one pass over the rows:
 *   d_lin = dout * a * mask;  d_x = dout * (1 - a);  d_alpha[t] = sum_{live rows of t} <dout, out - x> / a
 * with a = sigmoid(skip[t]) on rows whose row_gate != 0, 0 (passthrough) otherwise.  out = the forward's result, x = its
 * residual input, drop_mask [N, ldm] or NULL, d_alpha [T] is overwritten (d skip[t] = d_alpha[t] * a_t * (1 - a_t)). */
int wsi_skip_mix_bwd(const float* dout, int64_t ldd, const float* out, int64_t ldo, const float* x, int64_t ldx,
                     const float* drop_mask, int64_t ldm, const float* skip, const float* row_gate,
                     const int32_t* type_ptr_host, int T, int D, float* d_lin, int64_t ldl, float* d_x, int64_t lddx,
                     float* d_alpha, void* stream);

/* Weight gradient of the typed linear on tcgen05 (wgrad_tc.cu): dw[t] = dY_t^T . X_t, [T, M, Nn] fp32, overwritten.
 * Replaces autograd's per-type `grad_output.t() @ input` of every per-node-type nn.Linear (trainer/train_gnn.py:68-71
 * through models/HEATNet4.py:100-102,134,202).  dy_op / x_op are the WSI_OPF_BF16X3 operand forms ([2N, M] / [2N, Nn]
 * bf16: hi rows, then lo rows) the forward and data-gradient GEMMs already hold - nothing is transposed or re-converted;
 * product = hi.hi + hi.lo + lo.hi, fp32 accumulate.  M % 32 == 0, Nn % 8 == 0, N >= 512 (wsi_typed_wgrad_supported).
 * workspace: wsi_typed_wgrad_workspace_bytes(M, Nn, type_ptr_host, T) bytes (the per-chunk partial products). */
int wsi_typed_wgrad_supported(int64_t n_rows, int M, int Nn, int T);
int64_t wsi_typed_wgrad_workspace_bytes(int M, int Nn, const int32_t* type_ptr_host, int T);
int wsi_typed_wgrad(const void* dy_op, const void* x_op, int M, int Nn, const int32_t* type_ptr_host, int T, float* dw,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer step of the data-parallel training path (BASELINE config 5): torch.optim.Adam(lr, weight_decay) as the
 * reference builds it (parser.py:35-40; stepped by trainer/train_gnn.py:71) over ONE flat fp32 buffer of all
 * parameters (wsi_hgnn_b200/parallel.py keeps parameters, gradients and both moments flat).
 *   g' = grad * grad_scale + weight_decay * param;  exp_avg, exp_avg_sq updated in place;  step = 1, 2, ...;
 *   zero_grad != 0 also clears grad (the next step's zero_grad()).  All buffers 16 B aligned, n elements. */
int wsi_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, int zero_grad, void* stream);

/* The same step with torch.optim.Adam's treatment of parameters WITHOUT a gradient (skipped: no weight decay, no moment
 * decay, their own step count).  The flat buffer is a sequence of n_params slices [offs_dev[k], offs_dev[k+1]) (int64
 * [n_params + 1], 16 B aligned slices, offs_dev[n_params] == n); active_dev int32 [n_params] != 0 marks the parameters that
 * received a gradient this step (data-parallel: MAX-reduced across the ranks on the device - no host round trip);
 * steps_dev int32 [n_params] is incremented for the active ones; corr_ws: 8 * n_params bytes of scratch. */
int wsi_adam_step_masked(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int n_params,
                         const int64_t* offs_dev, const int32_t* active_dev, int32_t* steps_dev, float* corr_ws, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float grad_scale, int zero_grad,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WSI_HGNN_H_ */
