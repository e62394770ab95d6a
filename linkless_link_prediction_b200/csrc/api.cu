// Library-level entry points: version, error strings, device gate, launch counter.
#include "common.cuh"

namespace llp {

std::atomic<int64_t> g_launch_count{0};
int g_tuning[32] = {0};

long long* debug_buffer() {
  static long long* buf = nullptr;
  if (buf == nullptr) {
    if (cudaMalloc(&buf, 4096 * sizeof(long long)) != cudaSuccess) { cudaGetLastError(); buf = nullptr; return nullptr; }
    cudaMemset(buf, 0, 4096 * sizeof(long long));
  }
  return buf;
}

int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = 0;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  cached_dev = dev;
  cached_rc = (major == 10) ? 0 : LLP_E_DEVICE;
  return cached_rc;
}

}  // namespace llp

extern "C" int llp_version(void) { return 100; }

extern "C" int64_t llp_launch_count(void) { return llp::g_launch_count.load(); }

extern "C" int llp_device_supported(void) {
  int rc = llp::check_device();
  if (rc == 0) return 1;
  if (rc == LLP_E_DEVICE) return 0;
  return rc > 0 ? -rc - 1000 : rc;
}

extern "C" const char* llp_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case LLP_E_BADARG: return "llp: bad argument (null pointer, negative size or unknown enum)";
    case LLP_E_ALIGN: return "llp: pointer or leading dimension not aligned as required";
    case LLP_E_WORKSPACE: return "llp: workspace too small";
    case LLP_E_DEVICE: return "llp: current device is not an sm_100 (B200) GPU; there is no fallback path";
    case LLP_E_SHAPE: return "llp: shape not supported by the selected backend";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "llp: unknown error";
}

// development aid: copy the instrumentation scratch (see debug_buffer) to the host; synchronises the device
extern "C" int llp_debug_read(int64_t* host_out, int n) {
  if (host_out == nullptr || n < 0 || n > 4096) return LLP_E_BADARG;
  long long* buf = llp::debug_buffer();
  if (buf == nullptr) return LLP_E_DEVICE;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  e = cudaMemcpy(host_out, buf, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemset(buf, 0, 4096 * sizeof(long long));  // next instrumented launch starts clean
  return e == cudaSuccess ? 0 : (int)e;
}
