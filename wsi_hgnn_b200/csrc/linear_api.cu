// C-ABI front of the typed linear (include/wsi_hgnn.h): argument validation and dispatch between the
// tcgen05 tensor-core path (linear_tc.cu) and the fp32 SIMT path (linear_simt.cu).
#include "epilogue.cuh"

int wsi_typed_linear_simt_launch(const float* x, int64_t ldx, const float* w, int K, const int32_t* type_ptr_host,
                                 int T, const LinearEpilogue& ep, cudaStream_t stream);
// linear_tc.cu
bool wsi_typed_linear_tc_supported(int64_t n_rows, int K, int n_out, int64_t ldx);
bool wsi_opf_valid(int opf);
int64_t wsi_typed_linear_tc_workspace(int64_t n_rows, int K, int n_out, int T, int opf);
int wsi_typed_linear_tc_launch(const float* x, int64_t ldx, const float* w, int K, const int32_t* type_ptr_host,
                               int T, const LinearEpilogue& ep, int opf, void* workspace, int64_t workspace_bytes,
                               cudaStream_t stream);

int wsi_typed_linear_tc_gemm(const void* a_ws, const void* w_ws, int K, const int32_t* type_ptr_host, int T,
                             const LinearEpilogue& ep, void* y_op, int opf, cudaStream_t stream);
int wsi_split_launch(const float* a_src, int64_t a_ld, int64_t a_rows, void* a_dst, const float* b_src, int64_t b_ld,
                     int64_t b_rows, void* b_dst, int K, int opf, cudaStream_t stream, const int32_t* a_row_idx = nullptr);

extern "C" int wsi_typed_linear_tc_ok(int64_t n_rows, int K, int n_out) {
  return (wsi_typed_linear_tc_supported(n_rows, K, n_out, K) && n_out % 4 == 0) ? 1 : 0;
}

extern "C" int wsi_to_operand(const float* src, int64_t ld_src, int64_t rows, int K, int opf, void* dst, void* stream) {
  WSI_CHECK_ARG(rows >= 0 && K >= 8, "to_operand: bad shape");
  if (rows == 0) return WSI_OK;
  WSI_CHECK_ARG(src && dst && ld_src >= K, "to_operand: null pointer / short row stride");
  return wsi_split_launch(src, ld_src, rows, dst, nullptr, 4, 0, nullptr, K, opf, wsi_stream(stream));
}

// dst row i = operand form of src row row_idx[i]: the (dst, relation) segments of HGT gather their dst node's query
// (models/HGT.py:88-92 moved to the dst side) straight into the A operand of the relation-transform GEMM.
extern "C" int wsi_gather_to_operand(const float* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, int opf,
                                     void* dst, void* stream) {
  WSI_CHECK_ARG(rows >= 0 && K >= 8, "gather_to_operand: bad shape");
  if (rows == 0) return WSI_OK;
  WSI_CHECK_ARG(src && dst && row_idx && ld_src >= K, "gather_to_operand: null pointer / short row stride");
  return wsi_split_launch(src, ld_src, rows, dst, nullptr, 4, 0, nullptr, K, opf, wsi_stream(stream), row_idx);
}

int wsi_gather_rows16_launch(const void* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, void* dst,
                             cudaStream_t stream);
int wsi_convert_colsum_launch(const float* src, int64_t ld, int K, const int32_t* type_ptr_host, int T, void* dst,
                              float* colsum, float* partial, cudaStream_t stream);
int64_t wsi_convert_colsum_tiles(const int32_t* type_ptr_host, int T);

extern "C" int64_t wsi_to_operand_colsum_workspace_bytes(int K, const int32_t* type_ptr_host, int T) {
  if (!type_ptr_host || T < 1 || T > WSI_MAX_TYPES) return -1;
  return (wsi_convert_colsum_tiles(type_ptr_host, T) + 1) * (int64_t)K * 4;
}

// fp32 [N, K] -> WSI_OPF_BF16X3 operand form [2N, K] plus colsum[t, :] = sum of the rows of type t, one pass over src.
extern "C" int wsi_to_operand_colsum(const float* src, int64_t ld_src, int K, const int32_t* type_ptr_host, int T, void* dst,
                                     float* colsum, void* workspace, int64_t workspace_bytes, void* stream) {
  WSI_CHECK_ARG(type_ptr_host && T >= 1 && T <= WSI_MAX_TYPES, "to_operand_colsum: bad type_ptr / T=%d", T);
  WSI_CHECK_ARG(K >= 8 && K % 8 == 0 && ld_src >= K && ld_src % 4 == 0, "to_operand_colsum: K must be a multiple of 8, rows 16 B aligned");
  WSI_CHECK_ARG(src && dst && colsum && workspace && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dst) & 7) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                "to_operand_colsum: null / unaligned pointer");
  WSI_CHECK_ARG(workspace_bytes >= wsi_to_operand_colsum_workspace_bytes(K, type_ptr_host, T), "to_operand_colsum: workspace too small");
  return wsi_convert_colsum_launch(src, ld_src, K, type_ptr_host, T, dst, colsum, reinterpret_cast<float*>(workspace),
                                   wsi_stream(stream));
}

extern "C" int wsi_gather_rows16(const void* src, int64_t ld_src, const int32_t* row_idx, int64_t rows, int K, void* dst,
                                 void* stream) {
  WSI_CHECK_ARG(rows >= 0 && K >= 8 && K % 8 == 0 && ld_src >= K && ld_src % 8 == 0,
                "gather_rows16: K and the source row stride must be multiples of 8 elements");
  if (rows == 0) return WSI_OK;
  WSI_CHECK_ARG(src && dst && row_idx && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                "gather_rows16: null / unaligned pointer");
  return wsi_gather_rows16_launch(src, ld_src, row_idx, rows, K, dst, wsi_stream(stream));
}

extern "C" int wsi_typed_linear_op(const void* x_split, const void* w_split, const float* bias, int K, int n_out,
                                   const int32_t* type_ptr_host, int T, int act, const float* skip,
                                   const float* res, int64_t ldres, const float* drop_mask, int64_t ldmask,
                                   const float* row_gate, const float* row_scale, float* y, int64_t ldy,
                                   void* y_split, int opf, void* stream) {
  WSI_CHECK_ARG(type_ptr_host && T >= 1 && T <= WSI_MAX_TYPES, "typed_linear_op: bad type_ptr / T=%d", T);
  WSI_CHECK_ARG(act == WSI_ACT_NONE || act == WSI_ACT_GELU, "typed_linear_op: unknown activation %d", act);
  WSI_CHECK_ARG(!skip || res, "typed_linear_op: skip mix needs a residual");
  const int64_t n_rows = type_ptr_host[T];
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(x_split && w_split && (y || y_split), "typed_linear_op: null pointer");
  WSI_CHECK_ARG(!y || ldy >= n_out, "typed_linear_op: row stride smaller than the row");
  if (!wsi_typed_linear_tc_supported(n_rows, K, n_out, K) || n_out % 4 != 0) {
    wsi_set_error("typed_linear_op: shape (rows=%lld K=%d n_out=%d) does not fit the tcgen05 path",
                  (long long)n_rows, K, n_out);
    return WSI_ERR_UNSUPPORTED;
  }
  LinearEpilogue ep{};
  ep.bias = bias; ep.act = act; ep.skip = skip; ep.res = res; ep.ldres = ldres;
  ep.drop_mask = drop_mask; ep.ldmask = ldmask; ep.row_gate = row_gate; ep.row_scale = row_scale;
  ep.y = y; ep.ldy = ldy; ep.n_out = n_out;
  return wsi_typed_linear_tc_gemm(x_split, w_split, K, type_ptr_host, T, ep, y_split, opf, wsi_stream(stream));
}

extern "C" int64_t wsi_typed_linear_workspace_bytes(int64_t n_rows, int K, int n_out, int T, int impl, int opf) {
  if (impl == 1 || !wsi_opf_valid(opf)) return 0;
  if (!wsi_typed_linear_tc_supported(n_rows, K, n_out, K)) return 0;
  return wsi_typed_linear_tc_workspace(n_rows, K, n_out, T, opf);
}

extern "C" int wsi_typed_linear_f32(const float* x, int64_t ldx, const float* w, const float* bias, int K, int n_out,
                                    const int32_t* type_ptr_host, int T, int act, const float* skip,
                                    const float* res, int64_t ldres, const float* drop_mask, int64_t ldmask,
                                    const float* row_gate, const float* row_scale, float* y, int64_t ldy, int impl,
                                    int opf, void* workspace, int64_t workspace_bytes, void* stream) {
  WSI_CHECK_ARG(type_ptr_host && T >= 1 && T <= WSI_MAX_TYPES, "typed_linear: bad type_ptr / T=%d", T);
  WSI_CHECK_ARG(K >= 1 && n_out >= 1, "typed_linear: bad K=%d n_out=%d", K, n_out);
  WSI_CHECK_ARG(act == WSI_ACT_NONE || act == WSI_ACT_GELU, "typed_linear: unknown activation %d", act);
  WSI_CHECK_ARG(!skip || res, "typed_linear: skip mix needs a residual");
  WSI_CHECK_ARG(impl >= 0 && impl <= 2, "typed_linear: unknown impl %d", impl);
  WSI_CHECK_ARG(wsi_opf_valid(opf), "typed_linear: unknown operand format %d", opf);
  const int64_t n_rows = type_ptr_host[T];
  if (n_rows == 0) return WSI_OK;
  WSI_CHECK_ARG(x && w && y, "typed_linear: null pointer");
  WSI_CHECK_ARG(ldx >= K && ldy >= n_out, "typed_linear: row stride smaller than the row");
  LinearEpilogue ep{};
  ep.bias = bias; ep.act = act; ep.skip = skip; ep.res = res; ep.ldres = ldres;
  ep.drop_mask = drop_mask; ep.ldmask = ldmask; ep.row_gate = row_gate; ep.row_scale = row_scale;
  ep.y = y; ep.ldy = ldy; ep.n_out = n_out;
  const bool ptr_ok = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias) |
                        reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(drop_mask)) & 15) == 0 &&
                      ldy % 4 == 0 && (!res || ldres % 4 == 0) && (!drop_mask || ldmask % 4 == 0) &&
                      // a width that is not a multiple of 4 (the k-NN dot products against N candidates): the epilogue's
                      // last float4 spills into the row padding, so the pitch must cover it and there must be no
                      // per-column vector that would be read past its end
                      (n_out % 4 == 0 || (ldy >= ((n_out + 3) & ~3) && !bias && !res && !drop_mask));
  const bool tc_ok = wsi_typed_linear_tc_supported(n_rows, K, n_out, ldx) && ptr_ok && workspace != nullptr;
  if (impl == 2 && !tc_ok) {
    wsi_set_error("typed_linear: shape (rows=%lld K=%d n_out=%d ldx=%lld) does not fit the tcgen05 path",
                  (long long)n_rows, K, n_out, (long long)ldx);
    return WSI_ERR_UNSUPPORTED;
  }
  if (impl == 2 || (impl == 0 && tc_ok))
    return wsi_typed_linear_tc_launch(x, ldx, w, K, type_ptr_host, T, ep, opf, workspace, workspace_bytes,
                                      wsi_stream(stream));
  return wsi_typed_linear_simt_launch(x, ldx, w, K, type_ptr_host, T, ep, wsi_stream(stream));
}
