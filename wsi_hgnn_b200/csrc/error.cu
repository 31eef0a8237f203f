// Error channel and device queries of libwsi_hgnn.so (see include/wsi_hgnn.h "Conventions").
#include <stdlib.h>

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

bool wsi_pdl_enabled() { return getenv("WSI_NO_PDL") == nullptr; }

// Kernel-launch counter (bench.py reports it as `gpu_launches`): every WSI_CHECK_LAUNCH() bumps it.
static unsigned long long g_launches = 0;
void wsi_count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" int64_t wsi_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
