// cudaFuncSetAttribute configures a kernel on the CURRENT device only. A process may drive several
// devices (include/fmgpu.h: the device index of fmgpu_engine_create / fmgpu_channelizer_create;
// one host thread + streams per GPU), so "once" has to mean once per device:
// DeviceOnce runs a callable the first time it is reached on each device, thread-safe, and keeps
// that device's result.
#ifndef FMGPU_DEVICE_ONCE_H_
#define FMGPU_DEVICE_ONCE_H_

#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

namespace fmgpu {

class DeviceOnce {
 public:
  template <typename F>
  cudaError_t run(F &&f) {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0) {
      d = 0;
    }
    d %= kMaxDevices;
    if (done_[d].load(std::memory_order_acquire)) {
      return err_[d];
    }
    std::lock_guard<std::mutex> g(mu_);
    if (!done_[d].load(std::memory_order_relaxed)) {
      err_[d] = f();
      done_[d].store(true, std::memory_order_release);
    }
    return err_[d];
  }

 private:
  static constexpr int kMaxDevices = 64;
  std::mutex mu_;
  std::atomic<bool> done_[kMaxDevices] = {};
  cudaError_t err_[kMaxDevices] = {};
};

}  // namespace fmgpu

#endif  // FMGPU_DEVICE_ONCE_H_
