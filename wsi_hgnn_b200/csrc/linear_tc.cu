// tcgen05 typed linear - placeholder until the tensor-core kernel lands: reports "unsupported" so that
// impl=0 (auto) takes the SIMT path and impl=2 fails loudly.
#include "epilogue.cuh"

bool wsi_typed_linear_tc_supported(int64_t, int, int, int64_t) { return false; }
int64_t wsi_typed_linear_tc_workspace(int64_t, int, int, int) { return 0; }
int wsi_typed_linear_tc_launch(const float*, int64_t, const float*, int, const int32_t*, int, const LinearEpilogue&,
                               void*, int64_t, cudaStream_t) {
  wsi_set_error("typed_linear: tcgen05 path not built");
  return WSI_ERR_UNSUPPORTED;
}
