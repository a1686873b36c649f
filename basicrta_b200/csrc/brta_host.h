// Host-side helpers shared by the translation units of libbrta_gibbs.so.
#pragma once
#include <cuda_runtime.h>

namespace brta {

// record a message for brta_last_error() and return `code`
int fail(int code, const char* fmt, const char* detail = "");
int cuda_fail(cudaError_t e, const char* where);

}  // namespace brta
