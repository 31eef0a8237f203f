// Grouped (by node type, or by relation) linear, plain fp32 SIMT path.
//   Y[rows of group t] = X[rows of group t] . W_t^T  (+ fused epilogue, see epilogue.cuh)
// Serves the shapes the tcgen05 path does not take (K or n_out not tile aligned, e.g. HGT hidden 200,
// the [B, D] readout heads, the d_k x d_k relation transforms) and is the numerical cross-check of that path.
// Replaces the per-type nn.Linear calls of the reference (models/HEATNet4.py:100-102,134,202,219,243-245)
// and, with row indirection + per-head batching, the relation einsums of models/HGT.py:92-93.
#include "epilogue.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

struct SimtGemmArgs {
  const float* X; int64_t ldx;
  const float* W; int64_t w_group_stride;   // floats between W of consecutive groups
  int K;
  int w_kn;                 // 0: W_t is [n_out, K] (y = W x);  1: W_t is [K, n_out] (y = x W)
  const int* x_row_idx;     // optional gather of input rows  (indexed by the grouped row id)
  const int* y_row_idx;     // optional scatter of output rows
  int64_t x_z_stride, w_z_stride, y_z_stride;   // blockIdx.z batching (per-head block-diagonal transforms)
};

__global__ void __launch_bounds__(256)
typed_linear_simt_kernel(SimtGemmArgs a, TypeSegs segs, LinearEpilogue ep) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ int xrow[BM];
  const int tile_m = blockIdx.x;
  const int t = wsi_tile_group(segs, tile_m);
  const int row0 = segs.ptr[t] + (tile_m - segs.tile_start[t]) * BM;
  const int row_end = segs.ptr[t + 1];
  const int n0 = blockIdx.y * BN;
  const int Nout = ep.n_out;
  const int K = a.K;
  const float* X = a.X + (int64_t)blockIdx.z * a.x_z_stride;
  const float* Wt = a.W + (int64_t)t * a.w_group_stride + (int64_t)blockIdx.z * a.w_z_stride;

  const int tid = threadIdx.x;
  if (tid < BM) {
    int gr = row0 + tid;
    xrow[tid] = gr < row_end ? (a.x_row_idx ? __ldg(a.x_row_idx + gr) : gr) : -1;
  }
  __syncthreads();
  const int tx = tid % 16, ty = tid / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int r = idx / BK, c = idx % BK;
      int gk = k0 + c;
      int xr = xrow[r];
      As[c][r] = (xr >= 0 && gk < K) ? __ldg(X + (int64_t)xr * a.ldx + gk) : 0.f;
      if (!a.w_kn) {
        int gn = n0 + r;
        Bs[c][r] = (gn < Nout && gk < K) ? __ldg(Wt + (int64_t)gn * K + gk) : 0.f;
      }
    }
    if (a.w_kn) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int idx = tid + i * 256;
        int c = idx / BN, r = idx % BN;     // consecutive threads walk n (contiguous in a [K, n_out] matrix)
        int gk = k0 + c, gn = n0 + r;
        Bs[c][r] = (gn < Nout && gk < K) ? __ldg(Wt + (int64_t)gk * Nout + gn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const float alpha = ep.skip ? wsi_sigmoid(__ldg(ep.skip + t)) : 1.f;
  float* Y = ep.y + (int64_t)blockIdx.z * a.y_z_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gr = row0 + ty * TM + i;
    if (gr >= row_end) continue;
    int64_t row = a.y_row_idx ? __ldg(a.y_row_idx + gr) : gr;
    bool open = ep.row_gate ? (__ldg(ep.row_gate + row) != 0.f) : true;
    float rs = ep.row_scale ? __ldg(ep.row_scale + row) : 1.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= Nout) continue;
      Y[row * ep.ldy + n] = wsi_epilogue_value(ep, acc[i][j], t, row, n, alpha, open, rs);
    }
  }
}

// Few-row variant (the [T*B, D] readout and the [B, *] prediction heads at small batch, models/HEATNet4.py:219,243-245):
// one warp per (group, output column): the lanes stride over K with 16-byte loads of the W row and of up to RMAX
// activation rows, warp-shuffle reduction, fused epilogue by lane 0.  Launch-latency bound; the point is to spread
// the W read over many warps instead of one 64 x 64 tile block.
constexpr int RMAX = 8;

__global__ void __launch_bounds__(128)
typed_linear_fewrows_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int K,
                            TypeSegs segs, LinearEpilogue ep, int vec4) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int t = blockIdx.y;
  if (n >= ep.n_out) return;
  const int r0 = segs.ptr[t], r1 = segs.ptr[t + 1];
  if (r1 <= r0) return;
  const float* w = W + ((int64_t)t * ep.n_out + n) * K;
  const float alpha = ep.skip ? wsi_sigmoid(__ldg(ep.skip + t)) : 1.f;
  for (int rb = r0; rb < r1; rb += RMAX) {
    const int nr = min(RMAX, r1 - rb);
    float acc[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[r] = 0.f;
    if (vec4) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (r < nr) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(X + (int64_t)(rb + r) * ldx + k));
            acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
          }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int r = 0; r < RMAX; ++r)
          if (r < nr) acc[r] = fmaf(wv, __ldg(X + (int64_t)(rb + r) * ldx + k), acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r < nr) {
        const int64_t row = rb + r;
        const bool open = ep.row_gate ? (__ldg(ep.row_gate + row) != 0.f) : true;
        const float rs = ep.row_scale ? __ldg(ep.row_scale + row) : 1.f;
        ep.y[row * ep.ldy + n] = wsi_epilogue_value(ep, v, t, row, n, alpha, open, rs);
      }
    }
  }
}

int launch(const SimtGemmArgs& a, const int32_t* group_ptr_host, int T, const LinearEpilogue& ep, int batch,
           cudaStream_t stream) {
  TypeSegs segs;
  if (wsi_make_segs(&segs, group_ptr_host, T, BM) != 0) {
    wsi_set_error("grouped linear: bad group_ptr (T=%d, max %d)", T, WSI_MAX_TYPES);
    return WSI_ERR_ARG;
  }
  int tiles_m = segs.tile_start[T];
  if (tiles_m == 0 || ep.n_out == 0 || batch == 0) return WSI_OK;
  dim3 grid(tiles_m, (ep.n_out + BN - 1) / BN, batch);
  typed_linear_simt_kernel<<<grid, 256, 0, stream>>>(a, segs, ep);
  WSI_CHECK_LAUNCH();
  return WSI_OK;
}

}  // namespace

int wsi_typed_linear_simt_launch(const float* x, int64_t ldx, const float* w, int K, const int32_t* type_ptr_host,
                                 int T, const LinearEpilogue& ep, cudaStream_t stream) {
  int max_rows = 0;
  for (int t = 0; t < T; ++t) max_rows = max_rows > type_ptr_host[t + 1] - type_ptr_host[t] ? max_rows : type_ptr_host[t + 1] - type_ptr_host[t];
  if (max_rows <= 4 * RMAX && T <= 65535) {                // few rows per group: warp per output column
    TypeSegs segs;
    if (wsi_make_segs(&segs, type_ptr_host, T, 64) != 0) {
      wsi_set_error("typed_linear: bad type_ptr (T=%d, max %d)", T, WSI_MAX_TYPES);
      return WSI_ERR_ARG;
    }
    if (max_rows == 0) return WSI_OK;
    const int vec4 = K % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(w) & 15) == 0;
    dim3 grid((ep.n_out + 3) / 4, T);
    typed_linear_fewrows_kernel<<<grid, 128, 0, stream>>>(x, ldx, w, K, segs, ep, vec4);
    WSI_CHECK_LAUNCH();
    return WSI_OK;
  }
  SimtGemmArgs a{};
  a.X = x; a.ldx = ldx; a.W = w; a.w_group_stride = (int64_t)ep.n_out * K; a.K = K;
  return launch(a, type_ptr_host, T, ep, 1, stream);
}

extern "C" int wsi_rel_transform(const float* x, int64_t ldx, const int32_t* x_row_idx, const int32_t* y_row_idx,
                                 const float* w, const int32_t* rel_ptr_host, int R, int H, int d_k, int w_kn,
                                 float* y, int64_t ldy, void* stream) {
  WSI_CHECK_ARG(x && w && y && rel_ptr_host, "rel_transform: null pointer");
  WSI_CHECK_ARG(H >= 1 && d_k >= 1 && R >= 1, "rel_transform: bad H=%d d_k=%d R=%d", H, d_k, R);
  SimtGemmArgs a{};
  a.X = x; a.ldx = ldx; a.W = w; a.K = d_k; a.w_kn = w_kn;
  a.w_group_stride = (int64_t)H * d_k * d_k;
  a.x_row_idx = x_row_idx; a.y_row_idx = y_row_idx;
  a.x_z_stride = d_k; a.w_z_stride = (int64_t)d_k * d_k; a.y_z_stride = d_k;
  LinearEpilogue ep{};
  ep.y = y; ep.ldy = ldy; ep.n_out = d_k;
  return launch(a, rel_ptr_host, R, ep, H, wsi_stream(stream));
}
