// Internal helpers shared by the translation units of libpnb200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace pnbi {
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
void count_launch();
}  // namespace pnbi

#define PNBI_CUDA(call)                                       \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return pnbi::cuda_fail(e_, #call); \
  } while (0)
