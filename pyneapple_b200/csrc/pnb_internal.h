// Internal helpers shared by the translation units of libpnb200.so.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace pnbi {
int fail(int code, const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);
void count_launch();
// true for ordinary (pageable) host memory: cudaMemcpyAsync on it is staged by the driver at a
// fraction of PCIe speed and blocks the host, so the host pipelines stage it themselves
bool is_pageable(const void *ptr);
// memcpy split over host threads (half the cores, at most 12) (one thread does ~10 GB/s, PCIe 5 x16 moves ~55 GB/s)
void parallel_memcpy(void *dst, const void *src, size_t bytes);
// Makes `device` current for the lifetime of the object and restores the caller's device afterwards
// (the host entry points must not change the calling thread's current device).
class DeviceScope {
 public:
  explicit DeviceScope(int device) {
    if (cudaGetDevice(&prev_) != cudaSuccess) { cudaGetLastError(); prev_ = -1; }
    err_ = cudaSetDevice(device);
  }
  ~DeviceScope() {
    if (prev_ >= 0) cudaSetDevice(prev_);
  }
  cudaError_t error() const { return err_; }
  DeviceScope(const DeviceScope &) = delete;
  DeviceScope &operator=(const DeviceScope &) = delete;

 private:
  int prev_ = -1;
  cudaError_t err_ = cudaSuccess;
};
}  // namespace pnbi

#define PNBI_CUDA(call)                                       \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return pnbi::cuda_fail(e_, #call); \
  } while (0)
