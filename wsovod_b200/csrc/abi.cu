// abi.cu -- version / error strings / launch counter of libwsovod_b200.so
#include "common.cuh"

namespace wsovod {
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_tune[TUNE_COUNT] = {{0}, {1}, {1}};
}

WSOVOD_API int wsovod_b200_tune(int key, int value) {
  if (key < 0 || key >= wsovod::TUNE_COUNT) return WSOVOD_B200_EINVAL;
  return wsovod::g_tune[key].exchange(value, std::memory_order_relaxed);
}

WSOVOD_API int wsovod_b200_abi_version(void) { return WSOVOD_B200_ABI_VERSION; }

WSOVOD_API uint64_t wsovod_b200_launch_count(void) {
  return wsovod::g_launches.load(std::memory_order_relaxed);
}

WSOVOD_API const char* wsovod_b200_strerror(int code) {
  switch (code) {
    case 0: return "success";
    case WSOVOD_B200_EINVAL: return "invalid argument (null pointer, negative size or bad enum)";
    case WSOVOD_B200_ETOOBIG: return "a dimension exceeds what the kernels can index";
    case WSOVOD_B200_EWORKSPACE: return "workspace missing or too small (query *_workspace())";
    case WSOVOD_B200_EUNSUPPORTED: return "request not implemented by the sm_100a build";
    case WSOVOD_B200_EALIGN: return "pointer not 16-byte aligned";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown wsovod_b200 error";
}
