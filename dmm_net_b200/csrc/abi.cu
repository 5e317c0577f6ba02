// Library queries + error plumbing of the C ABI (include/dmm_b200.h).
#include "common.cuh"

namespace dmm {
static thread_local int g_last_cuda_error = 0;
void set_last_cuda_error(int e) { g_last_cuda_error = e; }
}  // namespace dmm

extern "C" int dmm_b200_version(void) { return DMM_B200_VERSION; }
extern "C" const char* dmm_b200_arch(void) { return "sm_100a"; }
extern "C" int dmm_b200_last_cuda_error(void) { return dmm::g_last_cuda_error; }

extern "C" const char* dmm_b200_error_string(int code) {
  switch (code) {
    case DMM_OK: return "ok";
    case DMM_ERR_INVALID_ARGUMENT: return "invalid argument";
    case DMM_ERR_UNSUPPORTED_SHAPE: return "shape beyond the compiled limits";
    case DMM_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case DMM_ERR_CUDA: return "CUDA runtime error (see dmm_b200_last_cuda_error)";
    default: return "unknown error";
  }
}

namespace dmm { int solver_max_rows(); int solver_max_cols(); }

extern "C" int dmm_b200_limits(int* out4) {
  if (!out4) return DMM_ERR_INVALID_ARGUMENT;
  out4[0] = dmm::solver_max_rows();
  out4[1] = dmm::solver_max_cols();
  out4[2] = 1 << 20;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
    out4[3] = sms;
  else {
    out4[3] = 0;
    cudaGetLastError();
  }
  return DMM_OK;
}
