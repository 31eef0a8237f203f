// fp32 -> operand form (WSI_OPF_*) stores shared by the conversion pre-pass, the GEMM epilogue and the row kernels.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint2 pack4_f16(float4 x) {
  const float lim = 65504.f;
  __half2 a = __floats2half2_rn(fminf(fmaxf(x.x, -lim), lim), fminf(fmaxf(x.y, -lim), lim));
  __half2 b = __floats2half2_rn(fminf(fmaxf(x.z, -lim), lim), fminf(fmaxf(x.w, -lim), lim));
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}
__device__ __forceinline__ uint2 pack4_bf16(float4 x) {
  __nv_bfloat162 a = __floats2bfloat162_rn(x.x, x.y), b = __floats2bfloat162_rn(x.z, x.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  return r;
}
// the operand-form store shared by this pre-pass, the GEMM epilogue and (hetero_attn.cu has its own copy) the attention
// kernel: 4 consecutive values of one row at `dst16` (16-bit element pointer)
template <int OPF>
__device__ __forceinline__ void store_operand4(void* dst16, int64_t lo_off, float4 x) {
  if (OPF == WSI_OPF_BF16X3) {
    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
    wsi_split_bf16(x.x, h0, l0); wsi_split_bf16(x.y, h1, l1); wsi_split_bf16(x.z, h2, l2); wsi_split_bf16(x.w, h3, l3);
    __nv_bfloat162 hv[2] = {__halves2bfloat162(h0, h1), __halves2bfloat162(h2, h3)};
    __nv_bfloat162 lv[2] = {__halves2bfloat162(l0, l1), __halves2bfloat162(l2, l3)};
    *reinterpret_cast<uint2*>(dst16) = *reinterpret_cast<uint2*>(hv);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst16) + lo_off) = *reinterpret_cast<uint2*>(lv);
  } else if (OPF == WSI_OPF_F16) {
    *reinterpret_cast<uint2*>(dst16) = pack4_f16(x);
  } else {
    *reinterpret_cast<uint2*>(dst16) = pack4_bf16(x);
  }
}

// run-time format: lo_off > 0 = WSI_OPF_BF16X3 (element offset of the lo plane), 0 = WSI_OPF_F16, < 0 = WSI_OPF_BF16
__device__ __forceinline__ void store_operand4_rt(void* dst16, int64_t lo_off, float4 x) {
  if (lo_off > 0) store_operand4<WSI_OPF_BF16X3>(dst16, lo_off, x);
  else if (lo_off == 0) store_operand4<WSI_OPF_F16>(dst16, 0, x);
  else store_operand4<WSI_OPF_BF16>(dst16, 0, x);
}

}  // namespace
