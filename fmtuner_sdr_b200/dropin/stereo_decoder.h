// Drop-in replacement for the reference's include/stereo_decoder.h (public surface of
// stereo_decoder.h:10-26), implemented over the C ABI in include/fmgpu.h.
#ifndef STEREO_DECODER_H
#define STEREO_DECODER_H

#include <stddef.h>
#include <stdint.h>

#include "dsp/liquid_primitives.h"

struct fmgpu_engine;

class StereoDecoder {
public:
  enum class BlendMode { Soft = 0, Normal = 1, Aggressive = 2 };

  StereoDecoder(int inputRate, int outputRate);
  ~StereoDecoder();
  StereoDecoder(const StereoDecoder &) = delete;
  StereoDecoder &operator=(const StereoDecoder &) = delete;

  // MPX at the DSP rate -> low-passed left / right at the DSP rate; one call = one logical
  // block for the stereo-lock counters
  size_t processAudio(const float *mono, float *left, float *right, size_t numSamples);
  void reset();
  void setForceStereo(bool force);
  void setForceMono(bool force);
  void setBlendMode(BlendMode mode);
  int getPilotLevelTenthsKHz() const;
  bool isStereo() const;

private:
  fmgpu_engine *engine_;
};

#endif
