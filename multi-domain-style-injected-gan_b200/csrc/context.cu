// Library context: device binding, error reporting, driver entry point for TMA descriptors.
#include "common.h"
#include "igemm.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace msig {

static thread_local char g_err[512] = "";
constexpr int kMaxDevices = 64;
static int g_sm_counts[kMaxDevices] = {0};   // per device, filled by msig_init(device)
static EncodeTiledFn g_encode = nullptr;
static std::mutex g_init_mu;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int sm_count() {                              // SM count of the CURRENT device (0 before its msig_init)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 0;
  return g_sm_counts[dev];
}
bool context_ready() { return g_encode != nullptr; }

static int g_pdl = -1;                        // -1: not decided yet (environment MSIG_PDL, default off)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("MSIG_PDL");
    g_pdl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_pdl != 0;
}
void set_pdl(bool on) { g_pdl = on ? 1 : 0; }
EncodeTiledFn encode_tiled() { return g_encode; }

}  // namespace msig

using namespace msig;

extern "C" {

int msig_version(void) { return MSIG_VERSION; }
const char* msig_last_error(void) { return g_err; }
int msig_sm_count(void) { return sm_count(); }
long long msig_kernel_launches(void) { return igemm_kernel_launches(); }
int msig_debug_set_pdl(int on) {
  set_pdl(on != 0);
  return MSIG_OK;
}

int msig_init(int device) {
  std::lock_guard<std::mutex> lock(g_init_mu);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(MSIG_ERR_CUDA, "msig_init: no CUDA device (%s); this library has no CPU path",
                     cudaGetErrorString(e));
  if (device < 0 || device >= count || device >= kMaxDevices)
    return set_error(MSIG_ERR_ARG, "msig_init: bad device %d", device);
  // The caller's current device is left untouched: launches go to the stream argument's device.
  cudaDeviceProp prop;
  MSIG_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(MSIG_ERR_UNSUPPORTED,
                     "msig_init: device %d is sm_%d%d; this library is built for sm_100a only", device,
                     prop.major, prop.minor);
  g_sm_counts[device] = prop.multiProcessorCount;
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MSIG_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
      return set_error(MSIG_ERR_CUDA, "msig_init: cuTensorMapEncodeTiled not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  return MSIG_OK;
}

}  // extern "C"
