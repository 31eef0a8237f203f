// Shared helpers for the wsi_hgnn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/wsi_hgnn.h"

#define WSI_OK 0

// Row groups of a grouped GEMM: node types of a typed linear (T <= 6 in the reference configs) or
// relations of a relation-grouped transform (R <= 72).
#define WSI_MAX_TYPES 128

void wsi_set_error(const char* fmt, ...);

#define WSI_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      wsi_set_error(__VA_ARGS__);                \
      return WSI_ERR_ARG;                        \
    }                                            \
  } while (0)

#define WSI_CHECK_CUDA(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      wsi_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return WSI_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

void wsi_count_launch();
#define WSI_CHECK_LAUNCH()              \
  do {                                  \
    wsi_count_launch();                 \
    WSI_CHECK_CUDA(cudaGetLastError()); \
  } while (0)

// Row ranges of the groups of a packed [N, *] matrix, passed to kernels by value.
struct TypeSegs {
  int T;
  int ptr[WSI_MAX_TYPES + 1];        // row range of group t: [ptr[t], ptr[t+1])
  int tile_start[WSI_MAX_TYPES + 1]; // first m-tile of group t (prefix of ceil(n_t / BM))
};

static inline int wsi_make_segs(TypeSegs* s, const int32_t* type_ptr_host, int T, int BM) {
  if (T < 1 || T > WSI_MAX_TYPES) return -1;
  s->T = T;
  s->tile_start[0] = 0;
  for (int t = 0; t <= T; ++t) s->ptr[t] = type_ptr_host[t];
  for (int t = 0; t < T; ++t) {
    int n = s->ptr[t + 1] - s->ptr[t];
    if (n < 0) return -1;
    s->tile_start[t + 1] = s->tile_start[t] + (n + BM - 1) / BM;
  }
  return 0;
}

// group of m-tile `tile` (binary search over the tile prefix; empty groups are skipped)
__device__ __forceinline__ int wsi_tile_group(const TypeSegs& segs, int tile) {
  int lo = 0, hi = segs.T - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs.tile_start[mid] <= tile) lo = mid; else hi = mid - 1;
  }
  // tile_start is non-decreasing; several empty groups may share the value - take the one that owns the tile
  while (lo + 1 < segs.T && segs.tile_start[lo + 1] <= tile) ++lo;
  return lo;
}

__device__ __forceinline__ float wsi_gelu(float x) {   // exact erf form == torch F.gelu default
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float wsi_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// fp32 -> (hi, lo) bf16 pair with hi + lo ~= x to ~2^-17 relative.
__device__ __forceinline__ void wsi_split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

static inline cudaStream_t wsi_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL): a kernel launched through wsi_launch_pdl may start while the previous kernel of
// the stream is still draining - its blocks are scheduled as SMs free up and run their prologue (barrier init, TMEM
// allocation, loads of PLAN data that no kernel writes) - and must call wsi_pdl_wait() before it touches anything an
// earlier kernel produced.  wsi_pdl_trigger() lets the NEXT kernel's blocks be scheduled early in the same way.
// A forward is ~13 kernels of 10-45 us: launch latency + prologue + tail of each were ~15 % of the step.
__device__ __forceinline__ void wsi_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void wsi_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Development knobs (error.cu): process-wide, read ONCE from the environment when the library is loaded (WSI_TC_DEBUG,
// WSI_ATTN_DEBUG, WSI_ATTN_KERNEL=ring|pipe, WSI_ATTN_RING, WSI_ATTN_BLOCKS, WSI_ATTN_CAP, WSI_ATTN_SEPARATE_MERGE,
// WSI_ATTN_STATIC, WSI_NO_PDL) and settable through wsi_dev_set(); no launch path calls getenv().  Not product API.
struct WsiDev {
  int tc_debug;            // typed_linear_tc_kernel: see TcArgs::dbg
  int tc_no_tma_store;     // typed_linear_tc_kernel: plain epilogue through per-lane global stores instead of TMA stores
  int attn_debug;          // attention kernels: see AttnArgs::dbg
  int attn_kernel;         // 0 = chosen per launch (default), 1 = register path, 2 = TMA ring, 3 = TMA pipe
  int attn_ring;           // ring depth of the TMA kernels (0 = default)
  int attn_blocks;         // blocks-per-SM cap of the TMA kernels (0 = none)
  int attn_cap;            // blocks-per-SM cap of the grid (0 = none)
  int attn_separate_merge; // merge the hub rows in a second launch
  int attn_static;         // static round-robin instead of the device work queue
  int no_pdl;              // launch without programmatic stream serialization
  int stream_debug;        // wsi_stream_forward: bit 0 = copy only the first `depth` blobs, bit 1 = no plan / forward
};
WsiDev* wsi_dev();
static inline bool wsi_pdl_enabled() { return wsi_dev()->no_pdl == 0; }

template <typename... KArgs, typename... Args>
static inline cudaError_t wsi_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = wsi_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
