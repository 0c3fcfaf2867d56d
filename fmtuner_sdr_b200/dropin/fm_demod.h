// Drop-in replacement for the reference's include/fm_demod.h: same class name and public
// signatures (fm_demod.h:11-36 of the reference), implemented over the C ABI in
// include/fmgpu.h instead of liquid-dsp. No liquid/liquid.h is needed to compile callers.
#ifndef FM_DEMOD_H
#define FM_DEMOD_H

#include <complex>
#include <stddef.h>
#include <stdint.h>

#include "dsp/liquid_primitives.h"

struct fmgpu_engine;

class FMDemod {
public:
  enum class DspAgcMode { Off = 0, Fast = 1, Slow = 2 };

  FMDemod(int inputRate, int outputRate);
  ~FMDemod();
  FMDemod(const FMDemod &) = delete;
  FMDemod &operator=(const FMDemod &) = delete;

  // uint8 IQ at the DSP rate -> mono audio at the output rate
  void process(const uint8_t *iq, float *audio, size_t numSamples);
  void processComplex(const std::complex<float> *iq, float *audio, size_t numSamples);
  // discriminator output only
  void processNoDownsample(const uint8_t *iq, float *audio, size_t numSamples);
  // MPX out (may be null) and, if monoOut is non-null, mono audio; returns mono frames
  size_t processSplit(const uint8_t *iq, float *mpxOut, float *monoOut, size_t numSamples);
  size_t processSplitComplex(const std::complex<float> *iq, float *mpxOut, float *monoOut,
                             size_t numSamples);
  size_t downsampleAudio(const float *demod, float *audio, size_t numSamples);
  void reset();

  void setDeemphasis(int tau_us);
  void setDeviation(double deviation);
  void setBandwidthMode(int mode);
  void setBandwidthHz(int bwHz);
  void setW0BandwidthHz(int bwHz);
  void setDspAgcMode(DspAgcMode mode);
  bool isClipping() const;
  float getClippingRatio() const;

private:
  fmgpu_engine *engine_;
};

#endif
