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
// Masked variant: the flat buffer is a sequence of parameter slices [offs[k], offs[k+1]); torch.optim.Adam leaves a
// parameter whose .grad is None alone (no weight decay, no moment decay) and counts ITS steps.  active[k] != 0 marks the
// parameters that received a gradient this step (decided on the device: in data-parallel runs the flags are MAX-reduced
// across the ranks, no host round trip).  adam_prep_kernel bumps the step count of the active parameters and derives
// their bias corrections; adam_masked_kernel finds the slice of every 16-byte group by binary search (<= 8 probes of an
// L1-resident table against 28 bytes of HBM traffic per element).
__global__ void adam_prep_kernel(const int* __restrict__ active, int* __restrict__ steps, float2* __restrict__ corr, int n_params,
                                 double lr, double b1, double b2) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_params) return;
  if (active[k] == 0) { corr[k] = make_float2(0.f, 0.f); return; }
  const int t = ++steps[k];
  corr[k] = make_float2((float)(lr / (1.0 - pow(b1, (double)t))), (float)(1.0 / sqrt(1.0 - pow(b2, (double)t))));
}

__global__ void __launch_bounds__(256) adam_masked_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, int64_t n4, const int64_t* __restrict__ offs,
                                                          const float2* __restrict__ corr, int n_params, float b1, float b2,
                                                          float eps, float wd, float gscale, int zero_grad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const int64_t e = i * 4;
    int lo = 0, hi = n_params - 1;                      // slice k with offs[k] <= e < offs[k + 1]
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(offs + mid) <= e) lo = mid; else hi = mid - 1;
    }
    const float2 c = __ldg(corr + lo);
    if (c.x != 0.f) {
      float4 P = reinterpret_cast<float4*>(p)[i];
      const float4 G = reinterpret_cast<const float4*>(g)[i];
      float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
      float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gr = fmaf(wd, pp[k], gg[k] * gscale);
        mm[k] = fmaf(b1, mm[k], (1.f - b1) * gr);
        vv[k] = fmaf(b2, vv[k], (1.f - b2) * gr * gr);
        pp[k] -= c.x * mm[k] / (sqrtf(vv[k]) * c.y + eps);
      }
      reinterpret_cast<float4*>(p)[i] = P;
      reinterpret_cast<float4*>(m)[i] = M;
      reinterpret_cast<float4*>(v)[i] = V;
    }
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
}  // namespace

extern "C" int wsi_adam_step_masked(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int n_params,
                                    const int64_t* offs_dev, const int32_t* active_dev, int32_t* steps_dev, float* corr_ws,
                                    float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                    int zero_grad, void* stream) {
  WSI_CHECK_ARG(n >= 0 && n % 4 == 0 && n_params >= 1, "adam_step_masked: n must be a multiple of 4 (16 B aligned slices)");
  if (n == 0) return WSI_OK;
  WSI_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && offs_dev && active_dev && steps_dev && corr_ws,
                "adam_step_masked: null pointer");
  WSI_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                  reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 && (reinterpret_cast<uintptr_t>(corr_ws) & 7) == 0,
                "adam_step_masked: buffers must be 16 B aligned");
  const int sms = wsi_num_sms();
  if (sms <= 0) return WSI_ERR_CUDA;
  cudaStream_t st = wsi_stream(stream);
  adam_prep_kernel<<<(n_params + 127) / 128, 128, 0, st>>>(active_dev, steps_dev, reinterpret_cast<float2*>(corr_ws), n_params,
                                                          (double)lr, (double)beta1, (double)beta2);
  WSI_CHECK_LAUNCH();
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
  adam_masked_kernel<<<(int)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n4, offs_dev,
                                                  reinterpret_cast<const float2*>(corr_ws), n_params, beta1, beta2, eps,
                                                  weight_decay, grad_scale, zero_grad);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

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
