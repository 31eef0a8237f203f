// Error channel and device queries of libwsi_hgnn.so (see include/wsi_hgnn.h "Conventions").
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/wsi_hgnn.h"

static thread_local char g_err[512] = "";

void wsi_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* wsi_last_error(void) { return g_err; }
extern "C" int wsi_abi_version(void) { return WSI_ABI_VERSION; }

extern "C" int wsi_num_sms(void) {
  int dev = 0, n = 0;
  WSI_CHECK_CUDA(cudaGetDevice(&dev));
  WSI_CHECK_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}

// Make `device` current for the calling thread (one process per GPU: called once by the Python host).
extern "C" int wsi_set_device(int device) {
  WSI_CHECK_CUDA(cudaSetDevice(device));
  return WSI_OK;
}

// ---- development knobs: the environment is read once, here, when the library is loaded
namespace {
int env_int(const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; }
WsiDev make_dev() {
  WsiDev d{};
  d.tc_debug = env_int("WSI_TC_DEBUG");
  d.attn_debug = env_int("WSI_ATTN_DEBUG");
  if (const char* k = getenv("WSI_ATTN_KERNEL")) d.attn_kernel = !strcmp(k, "vec") ? 1 : !strcmp(k, "ring") ? 2 : !strcmp(k, "pipe") ? 3 : 0;
  d.attn_ring = env_int("WSI_ATTN_RING");
  d.attn_blocks = env_int("WSI_ATTN_BLOCKS");
  d.attn_cap = env_int("WSI_ATTN_CAP");
  d.attn_separate_merge = getenv("WSI_ATTN_SEPARATE_MERGE") != nullptr;
  d.attn_static = getenv("WSI_ATTN_STATIC") != nullptr;
  d.no_pdl = getenv("WSI_NO_PDL") != nullptr;
  return d;
}
WsiDev g_dev = make_dev();
}  // namespace
WsiDev* wsi_dev() { return &g_dev; }

extern "C" int wsi_dev_set(const char* key, int value) {
  WSI_CHECK_ARG(key, "dev_set: null key");
  struct { const char* k; int* p; } tab[] = {
      {"tc_debug", &g_dev.tc_debug}, {"tc_no_tma_store", &g_dev.tc_no_tma_store}, {"attn_debug", &g_dev.attn_debug}, {"attn_kernel", &g_dev.attn_kernel},
      {"attn_ring", &g_dev.attn_ring}, {"attn_blocks", &g_dev.attn_blocks}, {"attn_cap", &g_dev.attn_cap},
      {"attn_separate_merge", &g_dev.attn_separate_merge}, {"attn_static", &g_dev.attn_static}, {"no_pdl", &g_dev.no_pdl}, {"stream_debug", &g_dev.stream_debug}};
  for (auto& e : tab)
    if (!strcmp(e.k, key)) { __atomic_store_n(e.p, value, __ATOMIC_RELAXED); return WSI_OK; }
  wsi_set_error("dev_set: unknown knob '%s'", key);
  return WSI_ERR_ARG;
}

// Kernel-launch counter (bench.py reports it as `gpu_launches`): every WSI_CHECK_LAUNCH() bumps it.
static unsigned long long g_launches = 0;
void wsi_count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" int64_t wsi_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
