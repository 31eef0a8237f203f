// Optimizer step of the data-parallel training path (BASELINE config 5): torch.optim.Adam(lr, weight_decay) as the
// reference builds it (parser.py:35-40; stepped by trainer/train_gnn.py:71) on ONE flat fp32 parameter / gradient / state
// buffer - one launch per step instead of torch's multi_tensor_apply chain over ~100 parameter tensors.
//   g' = g * grad_scale + wd * p;  m = b1 m + (1 - b1) g';  v = b2 v + (1 - b2) g'^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (torch.optim.Adam, amsgrad off, maximize off)
// HBM bound: 7 floats moved per element (p, g, m, v in; p, m, v out).
#include "common.cuh"

namespace {
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n4, int64_t n, float lr_c1, float inv_sqrt_c2,
                                                        float b1, float b2, float eps, float wd, float gscale, int zero_grad,
                                                        float* __restrict__ g_mut) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = fmaf(wd, pp[k], gg[k] * gscale);
      mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);
      vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
      pp[k] -= lr_c1 * mm[k] / (sqrtf(vv[k]) * inv_sqrt_c2 + eps);
    }
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = M;
    reinterpret_cast<float4*>(v)[i] = V;
    if (zero_grad) reinterpret_cast<float4*>(g_mut)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // tail (n % 4 elements)
  for (int64_t i = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = fmaf(wd, p[i], g[i] * gscale);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gr), vi = fmaf(b2, v[i], (1.f - b2) * gr * gr);
    m[i] = mi; v[i] = vi;
    p[i] -= lr_c1 * mi / (sqrtf(vi) * inv_sqrt_c2 + eps);
    if (zero_grad) g_mut[i] = 0.f;
  }
}
}  // namespace

extern "C" int wsi_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t step,
                             float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                             int zero_grad, void* stream) {
  WSI_CHECK_ARG(n >= 0 && step >= 1, "adam_step: bad n / step");
  if (n == 0) return WSI_OK;
  WSI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "adam_step: null pointer");
  WSI_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "adam_step: buffers must be 16 B aligned");
  const int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  const double c1 = 1.0 - pow((double)beta1, (double)step), c2 = 1.0 - pow((double)beta2, (double)step);
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  if (blocks < 1) blocks = 1;
  adam_flat_kernel<<<(int)blocks, 256, 0, wsi_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n4, n, (float)(lr / c1),
                                                                (float)(1.0 / sqrt(c2)), beta1, beta2, eps, weight_decay,
                                                                grad_scale, zero_grad, grad);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}
