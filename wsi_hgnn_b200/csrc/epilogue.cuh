// Fused epilogue shared by the SIMT and tcgen05 typed-linear kernels.
//   v = acc + bias[t][n]
//   v = gelu(v)                                   (HGT input projection, reference models/HGT.py:180)
//   v = v * drop_mask[row][n]                     (nn.Dropout on the a_linear output, models/HEATNet4.py:134)
//   v = v * sigmoid(skip[t]) + res[row][n] * (1 - sigmoid(skip[t]))   (models/HEATNet4.py:135)
//   rows whose gate is 0 (no incoming relation: the KeyError passthrough, models/HEATNet4.py:129-133) -> v = res
//   v = v * row_scale[row]                        (zero block for an empty node type, models/HEATNet4.py:240)
#pragma once
#include "common.cuh"

struct LinearEpilogue {
  const float* bias;        // [T, n_out] or nullptr
  int act;                  // WSI_ACT_*
  const float* skip;        // [T] or nullptr
  const float* res;         // [N, ldres] (required iff skip)
  int64_t ldres;
  const float* drop_mask;   // [N, ldmask] or nullptr
  int64_t ldmask;
  const float* row_gate;    // [N] or nullptr
  const float* row_scale;   // [N] or nullptr
  float* y;                 // [N, ldy]
  int64_t ldy;
  int n_out;
};

__device__ __forceinline__ float wsi_epilogue_value(const LinearEpilogue& ep, float acc, int t, int64_t row, int n,
                                                   float alpha, bool gate_open, float rscale) {
  float v = acc;
  if (ep.bias) v += __ldg(ep.bias + (int64_t)t * ep.n_out + n);
  if (ep.act == WSI_ACT_GELU) v = wsi_gelu(v);
  if (ep.drop_mask) v *= __ldg(ep.drop_mask + row * ep.ldmask + n);
  if (ep.skip) {
    float r = __ldg(ep.res + row * ep.ldres + n);
    v = gate_open ? (v * alpha + r * (1.0f - alpha)) : r;
  }
  return v * rscale;
}
