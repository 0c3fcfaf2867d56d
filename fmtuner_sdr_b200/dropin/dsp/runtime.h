// Drop-in replacement for the reference's include/dsp/runtime.h (runtime.h:11-30): block size
// plus the reset fan-out main.cpp registers. Instead of the reference's throw-away liquid
// self-test objects the constructor verifies that the CUDA engine library can reach a device.
#ifndef DSP_RUNTIME_H
#define DSP_RUNTIME_H

#include <cstddef>
#include <functional>
#include <vector>

namespace fm_tuner::dsp {

enum class ResetReason { Start = 0, Stop = 1, Retune = 2, ScanRestore = 3 };

const char *resetReasonName(ResetReason reason);

class Runtime {
public:
  Runtime(std::size_t blockSize, bool verbose);
  ~Runtime();

  std::size_t blockSize() const { return blockSize_; }
  void addResetHandler(std::function<void()> handler);
  void reset(ResetReason reason) const;

private:
  std::size_t blockSize_;
  bool verbose_;
  std::vector<std::function<void()>> handlers_;
};

}  // namespace fm_tuner::dsp

#endif
