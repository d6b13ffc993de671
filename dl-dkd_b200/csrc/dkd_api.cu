// Misc C-ABI entry points: version + error strings.
#include "dkd_common.cuh"

extern "C" int dkd_version(void) { return 100; }

extern "C" const char* dkd_error_string(int code) {
  switch (code) {
    case DKD_OK: return "ok";
    case DKD_ERR_ARG: return "dkd: bad argument (null pointer or negative size)";
    case DKD_ERR_SHAPE: return "dkd: unsupported shape";
    case DKD_ERR_ALIGN: return "dkd: pointer or leading dimension misaligned";
    case DKD_ERR_DRIVER: return "dkd: CUDA driver entry point unavailable";
    case DKD_ERR_WORKSPACE: return "dkd: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "dkd: unknown error";
}
