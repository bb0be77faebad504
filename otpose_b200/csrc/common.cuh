// Shared host/device helpers for the otpose_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/otpose_b200.h"

namespace otp {

void set_error(const char *fmt, ...);

inline int fail_arg(const char *what) {
  set_error("invalid argument: %s", what);
  return OTP_ERR_ARG;
}

inline int check_launch(const char *kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", kernel, cudaGetErrorString(e));
    return OTP_ERR_CUDA;
  }
  return OTP_OK;
}

#define OTP_REQUIRE(cond)                      \
  do {                                         \
    if (!(cond)) return ::otp::fail_arg(#cond); \
  } while (0)

constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
  return dev % kMaxDevices;
}

// per-device cache: a process may drive several GPUs (DataParallel, model.to("cuda:1"))
inline int num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] <= 0)
      sms[dev] = 148;  // B200
  }
  return sms[dev];
}

// cudaFuncSetAttribute applies to the CURRENT device only: one flag per device ordinal, not per process.
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool first() {
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// opt a kernel in to `bytes` of dynamic shared memory on the current device; false (+ otp_last_error) on failure
template <class F>
inline bool set_max_smem(F *kernel, size_t bytes, const char *name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu B dynamic shared memory): %s", name, bytes, cudaGetErrorString(e));
    return false;
  }
  return true;
}

// Kernel ids for launch counting and the optional per-kernel event timing
// (otp_profile_* in the C ABI).
enum KernelId {
  K_FINAL_PREDS = 0, K_MDCN, K_FUSION_SUM, K_FUSION_STACK, K_ADD_PE, K_UPSAMPLE, K_PYRAMID, K_CONV2D,
  K_BLOCK_FRONT, K_BLOCK_FOLD, K_BLOCK_APPLY, K_BLOCK_BACK, K_PACK, K_TC_FRONT, K_TC_APPLY, K_TC_BACK,
  K_TC_CONV, K_MDCN_BWD, K_FINAL_LAYER, K_CONV_BWD, K_FLOW_ENCODER, K_RSB, K_WINDOW, K_COUNT
};

// RAII around one (or n) kernel launch(es) on `st`: counts them and, when
// profiling is enabled, brackets them with CUDA events on the launching stream.
struct LaunchScope {
  LaunchScope(int id, cudaStream_t st, int nlaunch = 1);
  ~LaunchScope();
  int slot_;
  cudaStream_t st_;
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf-based GELU, the nn.GELU() default the reference MLP uses (model/blocks.py:250).
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

}  // namespace otp
