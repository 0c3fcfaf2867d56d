// Drop-in replacement for the reference's include/dsp/liquid_primitives.h, restricted to the
// type the reference's main.cpp names directly: fm_tuner::dsp::liquid::ComplexDecimator
// (liquid_primitives.h:161-188 of the reference). The other wrappers in that header (AGC,
// FIRFilter, NCO, FreqDemod, IIRFilterReal, Resampler) are implementation details of the
// reference's own DSP classes, which the engine replaces wholesale; they are not provided.
#ifndef DSP_LIQUID_PRIMITIVES_H
#define DSP_LIQUID_PRIMITIVES_H

#include <complex>
#include <cstddef>
#include <cstdint>

struct fmgpu_engine;

namespace fm_tuner::dsp::liquid {

class ComplexDecimator {
public:
  ComplexDecimator() = default;
  ~ComplexDecimator();
  ComplexDecimator(const ComplexDecimator &) = delete;
  ComplexDecimator &operator=(const ComplexDecimator &) = delete;

  void init(std::uint32_t factor, std::uint32_t tapsPerPhase = 12, float stopBandAtten = 70.0f);
  void reset();
  // uint8 IQ in, re-quantised uint8 IQ out
  std::size_t execute(const uint8_t *iqIn, std::size_t inSamples, uint8_t *iqOut,
                      std::size_t outCapacity) const;
  // uint8 IQ in, complex float out
  std::size_t executeComplex(const uint8_t *iqIn, std::size_t inSamples,
                             std::complex<float> *iqOut, std::size_t outCapacity) const;
  bool ready() const { return engine_ != nullptr || factor_ == 1; }
  std::uint32_t factor() const { return factor_; }

private:
  // mutable: a default-constructed decimator (factor 1, never initialised) converts as the
  // reference's does, and gets its engine on first use
  mutable fmgpu_engine *engine_ = nullptr;
  std::uint32_t factor_ = 1;
};

}  // namespace fm_tuner::dsp::liquid

#endif
