// Drop-in replacement for the reference's include/af_post_processor.h (public surface of
// af_post_processor.h:9-18), implemented over the C ABI in include/fmgpu.h.
#ifndef AF_POST_PROCESSOR_H
#define AF_POST_PROCESSOR_H

#include <stddef.h>
#include <stdint.h>

#include "dsp/liquid_primitives.h"

struct fmgpu_engine;

class AFPostProcessor {
public:
  AFPostProcessor(int inputRate, int outputRate);
  ~AFPostProcessor();
  AFPostProcessor(const AFPostProcessor &) = delete;
  AFPostProcessor &operator=(const AFPostProcessor &) = delete;

  void reset();
  void setDeemphasis(int tau_us);
  // DSP-rate left/right -> output-rate left/right; honours outCapacity like the reference loop
  size_t process(const float *inLeft, const float *inRight, size_t inSamples, float *outLeft,
                 float *outRight, size_t outCapacity);

private:
  fmgpu_engine *engine_;
};

#endif
