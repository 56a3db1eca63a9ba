// abi.cu -- library-level entry points: version, error reporting, device check.
#include "common.cuh"

#include <atomic>

namespace rn {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace rn

extern "C" int rn_abi_version(void) { return RN_ABI_VERSION; }

extern "C" const char* rn_last_error(void) { return rn::error_buffer(); }

namespace rn { unsigned long long launch_count(); }
extern "C" unsigned long long rn_launch_count(void) { return rn::launch_count(); }

extern "C" int rn_abi_struct_sizes(int32_t* out, int n) {
  const int32_t sizes[] = {(int32_t)sizeof(rn_relation_cfg), (int32_t)sizeof(rn_f_cfg),       (int32_t)sizeof(rn_conv_cfg),
                           (int32_t)sizeof(rn_conv_layer),   (int32_t)sizeof(rn_conv_grads),  (int32_t)sizeof(rn_lstm_cfg),
                           (int32_t)sizeof(rn_adam_cfg)};
  int i = 0;
  for (; out != nullptr && i < n && i < (int)(sizeof(sizes) / sizeof(sizes[0])); ++i) out[i] = sizes[i];
  return i;
}

extern "C" int rn_device_check(int device) {
  int major = 0, minor = 0;
  RN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  RN_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) return rn::fail(RN_ERR_ARCH, "device %d is sm_%d%d; librn_b200 is built for sm_100a only", device, major, minor);
  return RN_OK;
}
