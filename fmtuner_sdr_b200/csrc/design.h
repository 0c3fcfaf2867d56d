// design.h — host-side (design-time) filter synthesis for the engine.
//
// The reference designs every filter at construction through liquid-dsp:
//   liquid_firdes_kaiser        src/dsp/liquid_primitives.cpp:85-90,392-397
//   firfilt_crcf_create_kaiser  src/dsp/liquid_primitives.cpp:73-80
//   resamp_rrrf_create          src/dsp/liquid_primitives.cpp:338
//   symsync_crcf_create_rnyquist src/redsea_port/dsp/liquid_wrappers.cpp:153
// liquid-dsp is not a dependency of this engine; the published design formulas
// (Kaiser-windowed sinc, root-raised-cosine, polyphase partitioning) are
// evaluated here in double precision and rounded to float once.
#ifndef FMGPU_DESIGN_H_
#define FMGPU_DESIGN_H_

#include <cstdint>
#include <vector>

namespace fmdesign {

// Kaiser-windowed sinc low-pass prototype, n taps, cutoff fc (cycles/sample),
// stop-band attenuation As dB, fractional delay mu.
std::vector<float> kaiserLowpass(unsigned n, float fc, float As, float mu);

// Band-pass obtained by cosine-shifting a Kaiser low-pass to `center`, doubling
// it and normalising sum|h| to 1 (FIRFilter::init, liquid_primitives.cpp:83-112).
std::vector<float> shiftedBandpass(unsigned n, float fc, float As, float center);

// Root-raised-cosine pulse, k samples/symbol, m symbols delay: 2*k*m+1 taps.
std::vector<float> rootRaisedCosine(unsigned k, unsigned m, float beta);

// Arbitrary-rate resampler: polyphase bank [npfb][2m] laid out in WINDOW order
// (index 0 multiplies the oldest sample) and the 2^24 fixed-point phase step.
struct ResamplerDesign {
  unsigned npfb = 32;
  unsigned bits = 5;
  unsigned subLen = 0;
  uint32_t step = 1u << 24;
  std::vector<float> bank;
};
ResamplerDesign resampler(float rate, unsigned m, float fc, float As, unsigned npfb);
uint32_t resamplerStep(float rate);

// Symbol synchroniser banks (matched + derivative matched), [npfb][subLen] in
// window order, and its loop-filter constants.
struct SymSyncDesign {
  unsigned npfb = 32;
  unsigned subLen = 0;
  unsigned k = 3;
  std::vector<float> mf, dmf;
  float sosB0 = 0.0f, sosA1 = 0.0f, rateAdjustment = 0.0f;
};
SymSyncDesign symsyncRrc(unsigned k, unsigned m, float beta, unsigned npfb, float loopBandwidth);

// FMDemod::setBandwidthHz table lookup (fm_demod.cpp:99-135): returns the table
// index selected for bwHz given W0, and the filter parameters for that index.
struct ChannelFilterSpec {
  int index = 0;
  unsigned length = 81;
  float cutoff = 0.0f;
  float atten = 60.0f;
};
ChannelFilterSpec channelFilterSpec(int bwHz, int w0Hz, int inputRate);
int tefBandwidthHz(int mode);  // FMDemod::setBandwidthMode table, fm_demod.cpp:90-97

// NCO phase/frequency quantisation: radians -> 2^32 counts (nco_crcf, Appendix A.8)
uint32_t ncoConstrain(float radians);

}  // namespace fmdesign

#endif  // FMGPU_DESIGN_H_
