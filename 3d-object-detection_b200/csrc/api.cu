// api.cu -- version / error plumbing of libpp_b200.so.
#include "common.cuh"

namespace pp {
thread_local int g_last_cuda_error = 0;
}

extern "C" {

int pp_version(void) { return PP_B200_VERSION; }

int pp_last_cuda_error(void) { return pp::g_last_cuda_error; }

const char* pp_error_string(int code) {
  switch (code) {
    case PP_OK: return "ok";
    case PP_ERR_INVALID_ARG: return "invalid argument";
    case PP_ERR_WORKSPACE: return "workspace too small or misaligned";
    case PP_ERR_CUDA: return "CUDA runtime error (see pp_last_cuda_error)";
    case PP_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

}  // extern "C"
