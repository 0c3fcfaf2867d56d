// oracle/pipeline.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's block-processing classes on the
// IQ -> audio + RDS hot path, written against oracle/liquid_restated.hpp instead
// of liquid-dsp. Each class cites the reference file:line it follows
// (paths relative to /root/reference).
//
// PINNED to the reference's own code (round 2): the reference's UNMODIFIED
// fm_demod / stereo_decoder / af_post_processor / rds_decoder / liquid_primitives /
// redsea_port sources are compiled in place over oracle/liquid_shim into
// oracle/_ref/libfmref.so (oracle/Makefile), and tests/test_oracle_vs_reference.py
// demands identical floats, status, RDS bits and groups from this restatement
// (libm flavour) and that library: configs 1, 3, 5, the full 10 s of config 1 at
// both rates, resets, retunes, ragged class-level calls. Outputs of libfmref.so are
// also committed as fixtures (tests/golden/reference_*.npz). The integer RDS block
// synchroniser is pinned against block_sync.cpp compiled in place
// (tests/test_oracle_blocksync.py). What stays UNPINNED is below this file:
// liquid-dsp's own internals (liquid_restated.hpp), because no liquid build and no
// vectors exist in this image.
#ifndef ORACLE_PIPELINE_HPP_
#define ORACLE_PIPELINE_HPP_

#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

#include "liquid_restated.hpp"

namespace orc {

constexpr float kPiF = 3.14159265358979323846f;

// ===========================================================================
// fm_tuner::dsp::liquid::ComplexDecimator — src/dsp/liquid_primitives.cpp:364-499
// ===========================================================================
class ComplexDecimator {
public:
  void init(uint32_t factor, uint32_t tapsPerPhase = 12, float stopBandAtten = 70.0f) {
    if (factor == 0) {
      throw std::runtime_error("complex decimator factor must be >= 1");
    }
    factor_ = factor;
    tpp_ = std::max<uint32_t>(4, tapsPerPhase);
    block_.assign(factor_, cf32{});
    taps_.clear();
    ready_ = false;
    if (factor_ == 1) {
      return;
    }
    const uint32_t hLen = factor_ * tpp_;
    const float cutoff = std::clamp(0.45f / static_cast<float>(factor_), 0.01f, 0.45f);
    taps_ = firdes_kaiser(hLen, cutoff, stopBandAtten, 0.0f);
    dec_.create(factor_, taps_, 2.0f * cutoff);
    ready_ = true;
  }
  void reset() {
    if (factor_ != 1) {
      dec_.reset();
    }
  }
  // :461-499
  size_t executeComplex(const uint8_t *iqIn, size_t inSamples, cf32 *iqOut, size_t outCapacity) {
    if (!iqIn || !iqOut || inSamples == 0 || outCapacity == 0) {
      return 0;
    }
    static constexpr float kScale = 1.0f / 127.5f;
    if (factor_ == 1) {
      const size_t n = std::min(inSamples, outCapacity);
      for (size_t i = 0; i < n; i++) {
        iqOut[i].re = (static_cast<float>(iqIn[2 * i]) - 127.5f) * kScale;
        iqOut[i].im = (static_cast<float>(iqIn[2 * i + 1]) - 127.5f) * kScale;
      }
      return n;
    }
    if (!ready_) {
      return 0;
    }
    const size_t blocks = std::min(inSamples / factor_, outCapacity);
    for (size_t b = 0; b < blocks; b++) {
      const size_t inBase = b * factor_;
      for (size_t k = 0; k < factor_; k++) {
        const size_t idx = (inBase + k) * 2;
        block_[k].re = (static_cast<float>(iqIn[idx]) - 127.5f) * kScale;
        block_[k].im = (static_cast<float>(iqIn[idx + 1]) - 127.5f) * kScale;
      }
      iqOut[b] = dec_.execute(block_.data());
    }
    return blocks;
  }
  // :422-459 (uint8 re-quantised output)
  size_t execute(const uint8_t *iqIn, size_t inSamples, uint8_t *iqOut, size_t outCapacity) {
    if (!iqIn || !iqOut || inSamples == 0 || outCapacity == 0) {
      return 0;
    }
    if (factor_ == 1) {
      const size_t n = std::min(inSamples, outCapacity);
      std::copy_n(iqIn, n * 2, iqOut);
      return n;
    }
    if (!ready_) {
      return 0;
    }
    static constexpr float kScale = 1.0f / 127.5f;
    const size_t blocks = std::min(inSamples / factor_, outCapacity);
    for (size_t b = 0; b < blocks; b++) {
      const size_t inBase = b * factor_;
      for (size_t k = 0; k < factor_; k++) {
        const size_t idx = (inBase + k) * 2;
        block_[k].re = (static_cast<float>(iqIn[idx]) - 127.5f) * kScale;
        block_[k].im = (static_cast<float>(iqIn[idx + 1]) - 127.5f) * kScale;
      }
      const cf32 y = dec_.execute(block_.data());
      const float iOut = std::clamp((y.re * 127.5f) + 127.5f, 0.0f, 255.0f);
      const float qOut = std::clamp((y.im * 127.5f) + 127.5f, 0.0f, 255.0f);
      iqOut[2 * b] = static_cast<uint8_t>(iOut);
      iqOut[2 * b + 1] = static_cast<uint8_t>(qOut);
    }
    return blocks;
  }
  uint32_t factor() const { return factor_; }
  const std::vector<float> &taps() const { return taps_; }
  float scale() const { return dec_.scale(); }

private:
  uint32_t factor_ = 1;
  uint32_t tpp_ = 12;
  bool ready_ = false;
  std::vector<float> taps_;
  std::vector<cf32> block_;
  FirDecimC dec_;
};

// ===========================================================================
// FMDemod — src/fm_demod.cpp
// ===========================================================================
class FMDemod {
public:
  enum class DspAgcMode { Off = 0, Fast = 1, Slow = 2 };

  // :30-48
  FMDemod(int inputRate, int outputRate)
      : inputRate_(std::max(1, inputRate)), outputRate_(std::max(1, outputRate)) {
    const float iqCutoffNorm = std::clamp(110000.0f / static_cast<float>(inputRate_), 0.01f, 0.45f);
    iqFilter_.create_kaiser(81, iqCutoffNorm, 60.0f, 0.0f);
    dcI_.create_dc_blocker(0.0005f);
    dcQ_.create_dc_blocker(0.0005f);
    const float ratio = static_cast<float>(outputRate_) / static_cast<float>(inputRate_);
    if (ratio < 0.005f || ratio > 8.0f) {
      throw std::runtime_error("resampler ratio is out of supported range");
    }
    monoResampler_.create(ratio, 12, 0.47f, 60.0f, 32);
    monoDc_.create_dc_blocker(0.0008f);
    setDeviation(75000.0);
    setDeemphasis(75);
    setDspAgcMode(DspAgcMode::Off);
  }
  // :50-62
  void setDeemphasis(int tau_us) {
    if (tau_us <= 0) {
      deemphEnabled_ = false;
      return;
    }
    deemphEnabled_ = true;
    const float tau = static_cast<float>(tau_us) * 1e-6f;
    const float dt = 1.0f / static_cast<float>(outputRate_);
    const float alpha = dt / (tau + dt);
    monoDeemph_.create_b1_a2(alpha, 1.0f, -(1.0f - alpha));
  }
  // :64-71
  void setDeviation(double deviation) {
    deviation_ = deviation;
    freqdem_.create(static_cast<float>(deviation_ / static_cast<double>(inputRate_)));
  }
  // :73-88
  void reset() {
    clipping_ = false;
    clippingRatio_ = 0.0f;
    iqFilter_.reset();
    freqdem_.reset();
    dcI_.reset();
    dcQ_.reset();
    if (deemphEnabled_) {
      monoDeemph_.reset();
    }
    monoDc_.reset();
    monoResampler_.reset();
    if (agcReady_) {
      agc_.reset();
    }
  }
  // :90-97
  void setBandwidthMode(int mode) {
    static constexpr int kTefBwHz[] = {311000, 287000, 254000, 236000, 217000, 200000,
                                       184000, 168000, 151000, 133000, 114000, 97000,
                                       84000,  72000,  64000,  56000,  0};
    const int clipped = std::clamp(mode, 0, 16);
    setBandwidthHz(kTefBwHz[clipped]);
  }
  // :99-135
  void setBandwidthHz(int bwHz) {
    static constexpr std::array<int, 30> kXdrFmBwHz = {
        309000, 298000, 281000, 263000, 246000, 229000, 211000, 194000, 177000, 159000,
        142000, 125000, 108000, 95000,  90000,  83000,  73000,  63000,  55000,  48000,
        42000,  36000,  32000,  27000,  24000,  20000,  17000,  15000,  9000,   0};
    const int effectiveBwHz = (bwHz <= 0) ? w0BandwidthHz_ : bwHz;
    int selected = 29;
    if (effectiveBwHz > 0) {
      int minDiff = 0x7fffffff;
      for (int i = 0; i < 29; i++) {
        const int diff = std::abs(kXdrFmBwHz[static_cast<size_t>(i)] - effectiveBwHz);
        if (diff < minDiff) {
          minDiff = diff;
          selected = i;
        }
      }
    }
    if (selected == bandwidthMode_) {
      return;
    }
    bandwidthMode_ = selected;
    const int selectedBwHz = kXdrFmBwHz[static_cast<size_t>(selected)];
    const double nyquistHeadroomHz = 0.45 * static_cast<double>(inputRate_);
    const double iqCutoffHz =
        (selectedBwHz > 0)
            ? std::clamp(static_cast<double>(selectedBwHz) * 0.5, 9000.0, nyquistHeadroomHz)
            : nyquistHeadroomHz;
    const float cutoffNorm = std::clamp(
        static_cast<float>(iqCutoffHz / static_cast<double>(inputRate_)), 0.01f, 0.45f);
    const uint32_t filterLen = (selectedBwHz > 0 && selectedBwHz <= 73000) ? 121U : 81U;
    const float stopBandAtten = (selectedBwHz > 0 && selectedBwHz <= 42000) ? 70.0f : 60.0f;
    iqFilter_.create_kaiser(filterLen, cutoffNorm, stopBandAtten, 0.0f);
  }
  void setW0BandwidthHz(int bwHz) { w0BandwidthHz_ = std::clamp(bwHz, 0, 400000); }
  // :141-148
  void setDspAgcMode(DspAgcMode mode) {
    agcMode_ = mode;
    if (mode == DspAgcMode::Off) {
      return;
    }
    const float bandwidth = (mode == DspAgcMode::Fast) ? 0.01f : 0.001f;
    agc_.create(bandwidth, 1.0f);
    agcReady_ = true;
  }
  bool isClipping() const { return clipping_; }
  float getClippingRatio() const { return clippingRatio_; }

  // :150-180
  void demodulate(const uint8_t *iq, float *audio, size_t len) {
    size_t clipCount = 0;
    for (size_t i = 0; i < len; i++) {
      const uint8_t iByte = iq[2 * i];
      const uint8_t qByte = iq[2 * i + 1];
      if (iByte == 0 || iByte == 255 || qByte == 0 || qByte == 255) {
        clipCount++;
      }
      const float iRaw = (static_cast<float>(iByte) - 127.0f) / 127.5f;
      const float qRaw = (static_cast<float>(qByte) - 127.0f) / 127.5f;
      audio[i] = demodOne(iRaw, qRaw);
    }
    clipping_ = (clipCount > 0);
    clippingRatio_ = (len > 0) ? (static_cast<float>(clipCount) / static_cast<float>(len)) : 0.0f;
  }
  // :182-208
  void demodulateComplex(const cf32 *iq, float *audio, size_t len) {
    size_t clipCount = 0;
    for (size_t i = 0; i < len; i++) {
      const float iRaw = iq[i].re;
      const float qRaw = iq[i].im;
      if (std::fabs(iRaw) >= 0.995f || std::fabs(qRaw) >= 0.995f) {
        clipCount++;
      }
      audio[i] = demodOne(iRaw, qRaw);
    }
    clipping_ = (clipCount > 0);
    clippingRatio_ = (len > 0) ? (static_cast<float>(clipCount) / static_cast<float>(len)) : 0.0f;
  }
  // :210-226
  size_t downsampleAudio(const float *demod, float *audio, size_t numSamples) {
    size_t outCount = 0;
    float tmp[8];
    for (size_t i = 0; i < numSamples; i++) {
      const unsigned produced = std::min(8u, monoResampler_.execute(demod[i], tmp));
      for (unsigned p = 0; p < produced; p++) {
        float sample = tmp[p];
        if (deemphEnabled_) {
          sample = monoDeemph_.execute(sample);
        }
        sample = monoDc_.execute(sample);
        audio[outCount++] = sample;
      }
    }
    return outCount;
  }
  // :245-274
  size_t processSplit(const uint8_t *iq, float *mpxOut, float *monoOut, size_t numSamples) {
    scratch_.resize(std::max(scratch_.size(), numSamples));
    demodulate(iq, scratch_.data(), numSamples);
    if (mpxOut) {
      std::memcpy(mpxOut, scratch_.data(), numSamples * sizeof(float));
    }
    if (!monoOut) {
      return 0;
    }
    return downsampleAudio(scratch_.data(), monoOut, numSamples);
  }
  size_t processSplitComplex(const cf32 *iq, float *mpxOut, float *monoOut, size_t numSamples) {
    scratch_.resize(std::max(scratch_.size(), numSamples));
    demodulateComplex(iq, scratch_.data(), numSamples);
    if (mpxOut) {
      std::memcpy(mpxOut, scratch_.data(), numSamples * sizeof(float));
    }
    if (!monoOut) {
      return 0;
    }
    return downsampleAudio(scratch_.data(), monoOut, numSamples);
  }
  const FirFiltC &iqFilter() const { return iqFilter_; }
  float freqdemRef() const { return freqdem_.ref(); }

private:
  float demodOne(float iRaw, float qRaw) {
    const float iDc = dcI_.execute(iRaw);
    const float qDc = dcQ_.execute(qRaw);
    iqFilter_.push(cf32{iDc, qDc});
    cf32 y = iqFilter_.execute();
    if (agcMode_ != DspAgcMode::Off) {
      y = agc_.execute(y);
    }
    return freqdem_.demodulate(y);
  }

  int inputRate_;
  int outputRate_;
  double deviation_ = 75000.0;
  bool deemphEnabled_ = true;
  int bandwidthMode_ = 0;
  int w0BandwidthHz_ = 194000;
  DspAgcMode agcMode_ = DspAgcMode::Off;
  bool agcReady_ = false;
  bool clipping_ = false;
  float clippingRatio_ = 0.0f;
  std::vector<float> scratch_;
  FirFiltC iqFilter_;
  FreqDem freqdem_;
  Iir1 dcI_, dcQ_, monoDeemph_, monoDc_;
  Resamp monoResampler_;
  Agc agc_;
};

// ===========================================================================
// StereoDecoder — src/stereo_decoder.cpp
// ===========================================================================
class StereoDecoder {
public:
  enum class BlendMode { Soft = 0, Normal = 1, Aggressive = 2 };

  // :25-63
  StereoDecoder(int inputRate, int /*outputRate*/)
      : inputRate_(inputRate),
        pllFreq_(2.0f * kPiF * 19000.0f / static_cast<float>(inputRate)),
        pllMinFreq_(2.0f * kPiF * 18750.0f / static_cast<float>(inputRate)),
        pllMaxFreq_(2.0f * kPiF * 19250.0f / static_cast<float>(inputRate)) {
    int pilotTapCount =
        static_cast<int>(std::ceil(3.8 * static_cast<double>(inputRate_) / 3000.0));
    pilotTapCount = std::clamp(pilotTapCount, 63, 511);
    if ((pilotTapCount % 2) == 0) {
      pilotTapCount++;
    }
    const float pilotCenterNorm =
        std::clamp(19000.0f / static_cast<float>(inputRate_), 0.001f, 0.49f);
    const float pilotCutoffNorm =
        std::clamp(250.0f / static_cast<float>(inputRate_), 0.0005f, 0.45f);
    pilotFilter_.create(designShifted(static_cast<unsigned>(pilotTapCount), pilotCutoffNorm, 60.0f,
                                      pilotCenterNorm),
                        1.0f);
    const float audioCutoffNorm =
        std::clamp(15000.0f / static_cast<float>(inputRate_), 0.01f, 0.45f);
    leftFilter_.create_kaiser(121, audioCutoffNorm, 60.0f, 0.0f);
    rightFilter_.create_kaiser(121, audioCutoffNorm, 60.0f, 0.0f);
    delaySamples_ = std::max(0, (pilotTapCount - 1) / 2);
    delayLine_.assign(static_cast<size_t>(std::max(1, delaySamples_ + 1)), 0.0f);
    const float nominalPllFreq = 2.0f * kPiF * 19000.0f / static_cast<float>(inputRate_);
    pll_.create(nominalPllFreq);
    pll_.pll_set_bandwidth(0.01f);
  }

  // FIRFilter::init with center != 0 — src/dsp/liquid_primitives.cpp:83-112
  static std::vector<float> designShifted(unsigned length, float cutoff, float As, float center) {
    std::vector<float> taps = firdes_kaiser(length, cutoff, As, 0.0f);
    const int mid = static_cast<int>(length / 2);
    constexpr float kTwoPi = 6.28318530717958647692f;
    for (unsigned n = 0; n < length; n++) {
      const float phase = kTwoPi * center * static_cast<float>(static_cast<int>(n) - mid);
      taps[n] = 2.0f * taps[n] * std::cos(phase);  // design time: libm in both oracle builds
    }
    double sumAbs = 0.0;
    for (float tap : taps) {
      sumAbs += std::abs(tap);
    }
    if (sumAbs > 1e-12) {
      const float inv = static_cast<float>(1.0 / sumAbs);
      for (float &tap : taps) {
        tap *= inv;
      }
    }
    return taps;
  }

  // :67-86
  void reset() {
    stereoDetected_ = false;
    pilotMagnitude_ = 0.0f;
    pilotBandMagnitude_ = 0.0f;
    mpxMagnitude_ = 0.0f;
    stereoBlend_ = 0.0f;
    pilotLevelTenthsKHz_ = 0;
    pilotI_ = 0.0f;
    pilotQ_ = 0.0f;
    pllPhase_ = 0.0f;
    pllFreq_ = 2.0f * kPiF * 19000.0f / static_cast<float>(inputRate_);
    pilotCount_ = 0;
    pilotLossCount_ = 0;
    delayPos_ = 0;
    std::fill(delayLine_.begin(), delayLine_.end(), 0.0f);
    pilotFilter_.reset();
    pll_.reset();
    leftFilter_.reset();
    rightFilter_.reset();
  }
  void setForceStereo(bool f) { forceStereo_ = f; }
  void setForceMono(bool f) { forceMono_ = f; }
  void setBlendMode(BlendMode mode) { blendMode_ = mode; }
  int getPilotLevelTenthsKHz() const { return pilotLevelTenthsKHz_; }
  bool isStereo() const { return stereoDetected_; }

  // :92-286
  size_t processAudio(const float *mono, float *left, float *right, size_t numSamples) {
    if (!mono || !left || !right || numSamples == 0) {
      return 0;
    }
    constexpr float kMatrixScale = 0.5f;
    constexpr float kPilotRatioAcquire = 0.040f, kPilotRatioHold = 0.022f;
    constexpr float kMpxMinAcquire = 0.005f, kMpxMinHold = 0.0028f;
    constexpr float kPilotCoherenceAcquire = 0.18f, kPilotCoherenceHold = 0.11f;
    constexpr float kPllLockAcquireHz = 180.0f, kPllLockHoldHz = 320.0f;
    constexpr float kPilotEnvSmooth = 0.9995f;
    constexpr float kPilotEnvInject = 1.0f - kPilotEnvSmooth;
    constexpr float kPilotIqSmooth = 0.9995f;
    constexpr float kPilotIqInject = 1.0f - kPilotIqSmooth;

    float attackTau = 0.120f, releaseTau = 0.030f, lowQualityGate = 0.85f, lockFloor = 0.00f;
    if (blendMode_ == BlendMode::Soft) {
      attackTau = 0.090f;
      releaseTau = 0.040f;
      lowQualityGate = 0.75f;
    } else if (blendMode_ == BlendMode::Aggressive) {
      attackTau = 0.180f;
      releaseTau = 0.015f;
      lowQualityGate = 0.95f;
    }
    const float blendAttack =
        1.0f - std::exp(-1.0f / (attackTau * static_cast<float>(inputRate_)));
    const float blendRelease =
        1.0f - std::exp(-1.0f / (releaseTau * static_cast<float>(inputRate_)));
    const float nominalPllFreq = 2.0f * kPiF * 19000.0f / static_cast<float>(inputRate_);

    auto computeBlendTarget = [&](float pilotRatio, float pilotCoherence, float pllErrHz) -> float {
      if (forceMono_) {
        return 0.0f;
      }
      if (forceStereo_) {
        return 1.0f;
      }
      const float ratioQ =
          std::clamp((pilotRatio - kPilotRatioHold) /
                         std::max(kPilotRatioAcquire - kPilotRatioHold, 1e-4f),
                     0.0f, 1.0f);
      const float cohQ =
          std::clamp((pilotCoherence - kPilotCoherenceHold) /
                         std::max(kPilotCoherenceAcquire - kPilotCoherenceHold, 1e-4f),
                     0.0f, 1.0f);
      const float pllQ = std::clamp(
          (kPllLockHoldHz - pllErrHz) / std::max(kPllLockHoldHz - kPllLockAcquireHz, 1e-3f), 0.0f,
          1.0f);
      const float quality = std::min(ratioQ, std::min(cohQ, pllQ));
      float qualityShaped = quality * quality;
      if (blendMode_ == BlendMode::Soft) {
        qualityShaped = std::sqrt(std::max(0.0f, quality));
      } else if (blendMode_ == BlendMode::Aggressive) {
        qualityShaped = quality * quality * quality;
      }
      if (pilotRatio < (kPilotRatioHold * lowQualityGate) ||
          pilotCoherence < (kPilotCoherenceHold * lowQualityGate) ||
          pllErrHz > (kPllLockHoldHz * 1.10f)) {
        return 0.0f;
      }
      if (stereoDetected_) {
        return std::clamp(lockFloor + ((1.0f - lockFloor) * qualityShaped), 0.0f, 1.0f);
      }
      return 0.0f;
    };

    size_t outCount = 0;
    for (size_t i = 0; i < numSamples; i++) {
      const float mpx = mono[i];
      pilotFilter_.push(mpx);
      const float pilot = pilotFilter_.execute();
      pilotBandMagnitude_ =
          (pilotBandMagnitude_ * kPilotEnvSmooth) + (std::fabs(pilot) * kPilotEnvInject);
      mpxMagnitude_ = (mpxMagnitude_ * kPilotEnvSmooth) + (std::fabs(mpx) * kPilotEnvInject);
      const float phaseNow = pll_.phase();
      const float vcoI = m::cos(phaseNow);
      const float vcoQ = m::sin(phaseNow);
      const float error = pilot * vcoQ;
      pll_.pll_step(error);
      pll_.step();
      const float phaseNext = pll_.phase();
      float dphi = phaseNext - phaseNow;
      if (dphi > kPiF) {
        dphi -= 2.0f * kPiF;
      } else if (dphi < -kPiF) {
        dphi += 2.0f * kPiF;
      }
      pllPhase_ = phaseNext;
      pllFreq_ = std::clamp(dphi, pllMinFreq_, pllMaxFreq_);

      pilotI_ = (pilotI_ * kPilotIqSmooth) + ((pilot * vcoI) * kPilotIqInject);
      pilotQ_ = (pilotQ_ * kPilotIqSmooth) + ((pilot * vcoQ) * kPilotIqInject);
      const float pilotMagNow = std::sqrt((pilotI_ * pilotI_) + (pilotQ_ * pilotQ_));
      const float pilotRatioNow = pilotBandMagnitude_ / std::max(mpxMagnitude_, 1e-3f);
      const float pilotCoherenceNow = pilotMagNow / std::max(pilotBandMagnitude_, 1e-4f);
      const float pllErrHzNow = std::fabs(pllFreq_ - nominalPllFreq) *
                                static_cast<float>(inputRate_) / (2.0f * kPiF);
      const float targetStereoBlend =
          computeBlendTarget(pilotRatioNow, pilotCoherenceNow, pllErrHzNow);

      const float delayedMpx = delayLine_[delayPos_];
      delayLine_[delayPos_] = mpx;
      delayPos_++;
      if (delayPos_ >= delayLine_.size()) {
        delayPos_ = 0;
      }

      const float monoNorm = delayedMpx * kMatrixScale;
      const float pllRe = m::cos(pllPhase_);
      const float pllIm = m::sin(pllPhase_);
      const float cos2 = (pllRe * pllRe) - (pllIm * pllIm);
      const float lr = 2.0f * delayedMpx * cos2;
      const float stereoLeft = (delayedMpx + lr) * kMatrixScale;
      const float stereoRight = (delayedMpx - lr) * kMatrixScale;

      const float blendAlpha = (targetStereoBlend > stereoBlend_) ? blendAttack : blendRelease;
      stereoBlend_ += (targetStereoBlend - stereoBlend_) * blendAlpha;

      const float leftRaw = monoNorm + ((stereoLeft - monoNorm) * stereoBlend_);
      const float rightRaw = monoNorm + ((stereoRight - monoNorm) * stereoBlend_);

      leftFilter_.push(leftRaw);
      rightFilter_.push(rightRaw);
      left[outCount] = leftFilter_.execute();
      right[outCount] = rightFilter_.execute();
      outCount++;
    }

    // per-block tail :243-286
    const float pilotMag = std::sqrt((pilotI_ * pilotI_) + (pilotQ_ * pilotQ_));
    pilotMagnitude_ = (pilotMagnitude_ * 0.9f) + (pilotMag * 0.1f);
    const float mpxThreshold = stereoDetected_ ? kMpxMinHold : kMpxMinAcquire;
    const float pilotRatio = pilotBandMagnitude_ / std::max(mpxMagnitude_, 1e-3f);
    const float pilotCoherence = pilotMagnitude_ / std::max(pilotBandMagnitude_, 1e-4f);
    const float ratioThreshold = stereoDetected_ ? kPilotRatioHold : kPilotRatioAcquire;
    const float coherenceThreshold =
        stereoDetected_ ? kPilotCoherenceHold : kPilotCoherenceAcquire;
    const float pllErrHz =
        std::fabs(pllFreq_ - nominalPllFreq) * static_cast<float>(inputRate_) / (2.0f * kPiF);
    const float pllThreshold = stereoDetected_ ? kPllLockHoldHz : kPllLockAcquireHz;
    const bool pilotPresent = (mpxMagnitude_ > mpxThreshold) && (pilotRatio > ratioThreshold) &&
                              (pilotCoherence > coherenceThreshold) && (pllErrHz < pllThreshold);
    if (!forceStereo_) {
      if (!stereoDetected_) {
        if (pilotPresent) {
          pilotCount_++;
          pilotLossCount_ = 0;
          if (pilotCount_ >= 6) {
            stereoDetected_ = true;
          }
        } else {
          pilotCount_ = 0;
        }
      } else if (pilotPresent) {
        pilotLossCount_ = 0;
      } else if (++pilotLossCount_ >= 24) {
        stereoDetected_ = false;
        pilotCount_ = 0;
        pilotLossCount_ = 0;
      }
    }
    const float calibrated = pilotMagnitude_ * 8.0f;
    pilotLevelTenthsKHz_ =
        std::clamp(static_cast<int>(std::round(calibrated * 750.0f)), 0, 750);
    return outCount;
  }

  const FirFiltR &pilotFilter() const { return pilotFilter_; }
  const FirFiltR &audioFilter() const { return leftFilter_; }
  int delaySamples() const { return delaySamples_; }

private:
  int inputRate_;
  bool stereoDetected_ = false;
  bool forceStereo_ = false;
  bool forceMono_ = false;
  BlendMode blendMode_ = BlendMode::Normal;
  float pilotMagnitude_ = 0.0f;
  float pilotBandMagnitude_ = 0.0f;
  float mpxMagnitude_ = 0.0f;
  float stereoBlend_ = 0.0f;
  int pilotLevelTenthsKHz_ = 0;
  float pilotI_ = 0.0f;
  float pilotQ_ = 0.0f;
  float pllPhase_ = 0.0f;
  float pllFreq_;
  float pllMinFreq_;
  float pllMaxFreq_;
  int pilotCount_ = 0;
  int pilotLossCount_ = 0;
  std::vector<float> delayLine_;
  size_t delayPos_ = 0;
  int delaySamples_ = 0;
  FirFiltR pilotFilter_;
  Nco pll_;
  FirFiltR leftFilter_;
  FirFiltR rightFilter_;
};

// ===========================================================================
// AFPostProcessor — src/af_post_processor.cpp
// ===========================================================================
class AFPostProcessor {
public:
  // :7-18
  AFPostProcessor(int inputRate, int outputRate)
      : inputRate_(std::max(1, inputRate)), outputRate_(std::max(1, outputRate)) {
    const float ratio = static_cast<float>(outputRate_) / static_cast<float>(inputRate_);
    if (ratio < 0.005f || ratio > 8.0f) {
      throw std::runtime_error("resampler ratio is out of supported range");
    }
    leftResampler_.create(ratio, 12, 0.47f, 60.0f, 32);
    rightResampler_.create(ratio, 12, 0.47f, 60.0f, 32);
    leftDc_.create_dc_blocker(0.005f);
    rightDc_.create_dc_blocker(0.005f);
    reset();
    setDeemphasis(75);
  }
  // :20-29
  void reset() {
    leftResampler_.reset();
    rightResampler_.reset();
    leftDc_.reset();
    rightDc_.reset();
    if (deemphEnabled_) {
      leftDeemph_.reset();
      rightDeemph_.reset();
    }
  }
  // :31-45
  void setDeemphasis(int tau_us) {
    if (tau_us <= 0) {
      deemphEnabled_ = false;
      return;
    }
    deemphEnabled_ = true;
    const float tau = static_cast<float>(tau_us) * 1e-6f;
    const float samplePeriod = 1.0f / static_cast<float>(outputRate_);
    const float alpha = samplePeriod / (tau + samplePeriod);
    leftDeemph_.create_b1_a2(alpha, 1.0f, -(1.0f - alpha));
    rightDeemph_.create_b1_a2(alpha, 1.0f, -(1.0f - alpha));
  }
  // :47-78
  size_t process(const float *inLeft, const float *inRight, size_t inSamples, float *outLeft,
                 float *outRight, size_t outCapacity) {
    if (!inLeft || !inRight || !outLeft || !outRight || inSamples == 0 || outCapacity == 0) {
      return 0;
    }
    size_t outCount = 0;
    float lt[8], rt[8];
    for (size_t i = 0; i < inSamples && outCount < outCapacity; i++) {
      const unsigned lp = std::min(8u, leftResampler_.execute(inLeft[i], lt));
      const unsigned rp = std::min(8u, rightResampler_.execute(inRight[i], rt));
      const unsigned produced = std::min(lp, rp);
      for (unsigned idx = 0; idx < produced && outCount < outCapacity; idx++) {
        float l = lt[idx];
        float r = rt[idx];
        if (deemphEnabled_) {
          l = leftDeemph_.execute(l);
          r = rightDeemph_.execute(r);
        }
        l = leftDc_.execute(l);
        r = rightDc_.execute(r);
        outLeft[outCount] = l;
        outRight[outCount] = r;
        outCount++;
      }
    }
    return outCount;
  }
  const Resamp &resampler() const { return leftResampler_; }
  float deemphB0() const { return leftDeemph_.b0(); }
  float deemphA1() const { return leftDeemph_.a1(); }

private:
  int inputRate_;
  int outputRate_;
  bool deemphEnabled_ = false;
  Iir1 leftDeemph_, rightDeemph_, leftDc_, rightDc_;
  Resamp leftResampler_, rightResampler_;
};

// ===========================================================================
// RDS: redsea_port — src/redsea_port/dsp/subcarrier.cpp, liquid_wrappers.cpp,
//                     block_sync.cpp, group.cpp ; facade src/rds_decoder.cpp
// ===========================================================================
struct RDSGroup {  // include/rds_decoder.h:9-15
  uint16_t blockA, blockB, blockC, blockD;
  uint8_t errors;
};

// subcarrier.cpp:50-86
class BiphaseDecoder {
public:
  // returns has_value; *bit is the decision
  bool push(cf32 psk, bool *bit) {
    const float biphase_re = (psk.re - prev_.re) * 0.5f;
    *bit = biphase_re >= 0.0f;
    const bool has_value = (clock_ % 2 == clock_polarity_);
    prev_ = psk;
    clock_history_[clock_] = std::fabs(biphase_re);
    clock_++;
    if (clock_ == clock_history_.size()) {
      float even_sum = 0.0f, odd_sum = 0.0f;
      for (size_t i = 0; i < clock_history_.size(); i += 2) {
        even_sum += clock_history_[i];
        odd_sum += clock_history_[i + 1];
      }
      if (even_sum > odd_sum) {
        clock_polarity_ = 0;
      } else if (odd_sum > even_sum) {
        clock_polarity_ = 1;
      }
      clock_history_.fill(0.0f);
      clock_ = 0;
    }
    return has_value;
  }

private:
  cf32 prev_{};
  std::array<float, 128> clock_history_{};
  uint32_t clock_ = 0;
  uint32_t clock_polarity_ = 0;
};

// liquid_wrappers.cpp:98-147 — redsea's NCO wrapper (only data stream 0 is used, rds_decoder.cpp:89)
class RdsNco {
public:
  void init(float freq) {
    nco_.create(freq);
    prev_f0_phase_ = 0.0f;
    phase0_ = 0.0f;
  }
  void reset() { nco_.reset(); }  // prev_f0_phase_ / phases_ are NOT reset (:116-119)
  void setPLLBandwidth(float bw) { nco_.pll_set_bandwidth(bw); }
  void stepPLL(float dphi) { nco_.pll_step(dphi); }
  cf32 mixDown(float s) const {
    const float a = -phase0_;
    return cf32{s * m::cos(a), s * m::sin(a)};
  }
  void step() {
    nco_.step();
    const float phase_now = nco_.phase();
    const float delta = unwrap(phase_now - prev_f0_phase_);
    prev_f0_phase_ = phase_now;
    phase0_ = unwrap(phase0_ + ((delta * 57000.f) / 57000.f));
  }

private:
  static float unwrap(float p) {
    constexpr float k2Pi = 2.f * kPiF;
    if (p > kPiF) {
      return p - k2Pi;
    }
    if (p < -kPiF) {
      return p + k2Pi;
    }
    return p;
  }
  Nco nco_;
  float prev_f0_phase_ = 0.0f;
  float phase0_ = 0.0f;
};

// ---- integer back end: block_sync.cpp / group.cpp restated ------------------
enum class Offset : uint8_t { A = 0, B = 1, C = 2, Cprime = 3, D = 4, invalid = 5 };

inline uint32_t rdsSyndrome(uint32_t v) {  // block_sync.cpp:85-128
  static constexpr uint32_t H[26] = {
      0b1000000000, 0b0100000000, 0b0010000000, 0b0001000000, 0b0000100000, 0b0000010000,
      0b0000001000, 0b0000000100, 0b0000000010, 0b0000000001, 0b1011011100, 0b0101101110,
      0b0010110111, 0b1010000111, 0b1110011111, 0b1100010011, 0b1101010101, 0b1101110110,
      0b0110111011, 0b1000000001, 0b1111011100, 0b0111101110, 0b0011110111, 0b1010100111,
      0b1110001111, 0b1100011011};
  uint32_t r = 0;
  for (unsigned k = 0; k < 26; k++) {
    if ((v >> k) & 1u) {
      r ^= H[25 - k];
    }
  }
  return r;
}

inline Offset rdsOffsetForSyndrome(uint32_t s) {  // :71-81
  switch (s) {
  case 0b1111011000: return Offset::A;
  case 0b1111010100: return Offset::B;
  case 0b1001011100: return Offset::C;
  case 0b1111001100: return Offset::Cprime;
  case 0b1001011000: return Offset::D;
  default: return Offset::invalid;
  }
}

inline uint32_t rdsOffsetWord(Offset o) {  // :138-144
  static constexpr uint32_t W[5] = {0b0011111100, 0b0110011000, 0b0101101000, 0b1101010000,
                                    0b0110110100};
  return W[static_cast<int>(o)];
}

inline int rdsBlockNumber(Offset o) {  // :42-53
  switch (o) {
  case Offset::A: return 0;
  case Offset::B: return 1;
  case Offset::C:
  case Offset::Cprime: return 2;
  case Offset::D: return 3;
  default: return 0;
  }
}

inline Offset rdsNextOffset(Offset o) {  // :56-67
  switch (o) {
  case Offset::A: return Offset::B;
  case Offset::B: return Offset::C;
  case Offset::C:
  case Offset::Cprime: return Offset::D;
  case Offset::D: return Offset::A;
  default: return Offset::A;
  }
}

class BlockStream {
public:
  struct Blk {
    uint16_t data = 0;
    bool received = false;
    bool had_errors = false;
  };
  struct Grp {
    Blk b[4];
  };

  void pushBit(bool bit) {  // :254-264
    reg_ = (reg_ << 1) + (bit ? 1u : 0u);
    until_--;
    bitcount_++;
    if (until_ == 0) {
      findBlock();
      until_ = in_sync_ ? 26 : 1;
    }
  }
  bool hasGroupReady() const { return has_ready_; }
  Grp popGroup() {
    has_ready_ = false;
    return ready_;
  }

private:
  struct Pulse {
    Offset offset = Offset::invalid;
    uint32_t pos = 0;
  };
  static bool couldFollow(const Pulse &p, const Pulse &other) {  // :189-198
    const uint32_t d = p.pos - other.pos;
    return d % 26 == 0 && d / 26 <= 6 && p.offset != Offset::invalid &&
           other.offset != Offset::invalid &&
           (static_cast<uint32_t>(rdsBlockNumber(other.offset)) + d / 26) % 4 ==
               static_cast<uint32_t>(rdsBlockNumber(p.offset));
  }
  // correctBurstErrors :165-184 with the lookup order of makeErrorLookupTable :133-162
  static bool correctBurst(uint32_t raw, Offset expected, uint32_t *corrected) {
    const uint32_t syn = rdsSyndrome(raw);
    const uint32_t word = rdsOffsetWord(expected);
    for (uint32_t ebits : {0b1u, 0b11u}) {
      for (uint32_t shift = 0; shift < 26; shift++) {
        const uint32_t ev = (ebits << shift) & 0x3ffffffu;
        if (rdsSyndrome(ev ^ word) == syn) {
          *corrected = raw ^ ev;
          return true;
        }
      }
    }
    return false;
  }
  void findBlock() {  // :267-313
    const uint32_t raw = reg_ & 0x3ffffffu;
    Offset off = rdsOffsetForSyndrome(rdsSyndrome(raw));
    // acquireSync :235-251
    if (!in_sync_) {
      bits_since_lost_++;
      if (off != Offset::invalid) {
        for (int i = 0; i < 3; i++) {
          pulses_[i] = pulses_[i + 1];
        }
        pulses_[3] = Pulse{off, bitcount_};
        bool found = false;
        for (int i1 = 0; i1 < 2 && !found; i1++) {
          for (int i2 = i1 + 1; i2 < 3 && !found; i2++) {
            if (couldFollow(pulses_[3], pulses_[i2]) && couldFollow(pulses_[i2], pulses_[i1])) {
              found = true;
            }
          }
        }
        if (found) {
          in_sync_ = true;
          expected_ = off;
          cur_ = Grp{};
          bits_since_lost_ = 0;
        }
      }
    }
    if (!in_sync_) {
      return;
    }
    if (expected_ == Offset::C && off == Offset::Cprime) {
      expected_ = Offset::Cprime;
    }
    const bool had_errors = (off != expected_);
    err50_[err_ptr_] = had_errors ? 1 : 0;
    err_ptr_ = (err_ptr_ + 1) % 50;
    int sum = 0;
    for (int e : err50_) {
      sum += e;
    }
    if (sum > 42) {
      in_sync_ = false;
      err50_.fill(0);
      return;
    }
    uint16_t data = static_cast<uint16_t>(raw >> 10);
    if (had_errors) {  // use_fec = true (rds_decoder.cpp:18)
      uint32_t corrected = 0;
      if (correctBurst(raw, expected_, &corrected)) {
        data = static_cast<uint16_t>(corrected >> 10);
        off = expected_;
      }
    }
    if (off == expected_) {
      Blk &blk = cur_.b[rdsBlockNumber(expected_)];
      blk.data = data;
      blk.received = true;
      blk.had_errors = had_errors;
    }
    const Offset next = rdsNextOffset(expected_);
    if (next == Offset::A) {
      ready_ = cur_;
      has_ready_ = true;
      cur_ = Grp{};
    }
    expected_ = next;
  }

  uint32_t bitcount_ = 0;
  uint32_t until_ = 1;
  uint32_t reg_ = 0;
  Offset expected_ = Offset::A;
  bool in_sync_ = false;
  std::array<int, 50> err50_{};
  int err_ptr_ = 0;
  Grp cur_{}, ready_{};
  bool has_ready_ = false;
  uint32_t bits_since_lost_ = 0;
  Pulse pulses_[4];
};

// rds_decoder.cpp:29-58
inline RDSGroup packGroup(const BlockStream::Grp &g) {
  auto val = [&](int i) -> uint16_t { return g.b[i].received ? g.b[i].data : 0; };
  auto err = [&](int i) -> uint8_t {
    if (!g.b[i].received) {
      return 3;
    }
    return g.b[i].had_errors ? 1 : 0;
  };
  return RDSGroup{val(0), val(1), val(2), val(3),
                  static_cast<uint8_t>((err(0) << 6) | (err(1) << 4) | (err(2) << 2) | err(3))};
}

// subcarrier.cpp:94-235 (single data stream)
class SubcarrierSet {
public:
  static constexpr float kTargetRate = 171000.f;
  explicit SubcarrierSet(float samplerate) : resample_ratio_(kTargetRate / samplerate) {
    resampler_.create(1.f, 13, 0.47f, 60.0f, 32);
    agc_.create(500.0f / kTargetRate, 0.08f);
    lpf_.create_kaiser(255, 2400.0f / kTargetRate, 60.0f, 0.0f);
    symsync_.create_rnyquist_rrc(3, 3, 0.8f, 32);
    symsync_.set_lf_bw(2200.0f / kTargetRate);
    osc_.init(57000.f * (2.f * kPiF) / kTargetRate);
    osc_.setPLLBandwidth(0.03f / kTargetRate);
    if (resample_ratio_ < 0.005f || resample_ratio_ > 2.0f) {
      throw std::runtime_error("error: Can't support this sample rate");
    }
    resampler_.set_rate(resample_ratio_);
  }
  // :108-114
  void reset() {
    symsync_.reset();
    osc_.reset();
    sample_num_since_reset_ = 0;
  }
  // :153-235 ; appends decoded bits
  void chunkToBits(const float *mpx, size_t n, std::vector<uint8_t> &bits) {
    float rs[2];
    for (size_t i = 0; i < n; i++) {
      unsigned nr;
      if (resample_ratio_ == 1.0f) {
        rs[0] = mpx[i];
        nr = 1;
      } else {
        nr = resampler_.execute(mpx[i], rs);
      }
      for (unsigned j = 0; j < nr; j++) {
        sample171(rs[j], bits);
      }
    }
  }
  const Resamp &resampler() const { return resampler_; }
  const FirFiltC &lpf() const { return lpf_; }
  const SymSync &symsync() const { return symsync_; }

private:
  void sample171(float s, std::vector<uint8_t> &bits) {
    const cf32 bb = osc_.mixDown(s);
    lpf_.push(bb);
    if (sample_num_since_reset_ % 24 == 0) {
      const cf32 lo = agc_.execute(lpf_.execute());
      cf32 out[8];
      const unsigned n_out = symsync_.step(lo, out);
      if (n_out == 1) {
        const cf32 sym = out[0];
        const float phase_error = std::clamp(bpsk_phase_error(sym), -kPiF, kPiF);
        osc_.stepPLL(phase_error * 12.0f);
        bool b;
        if (biphase_.push(sym, &b)) {
          const bool bit = (b != delta_prev_);
          delta_prev_ = b;
          bits.push_back(bit ? 1 : 0);
        }
      }
    }
    osc_.step();
    sample_num_since_reset_++;
  }

  const float resample_ratio_;
  uint32_t sample_num_since_reset_ = 0;
  Resamp resampler_;
  Agc agc_;
  FirFiltC lpf_;
  SymSync symsync_;
  RdsNco osc_;
  BiphaseDecoder biphase_;
  bool delta_prev_ = false;
};

// rds_decoder.cpp
class RDSDecoder {
public:
  explicit RDSDecoder(int inputRate) : rate_(std::max(1, inputRate)), sub_(static_cast<float>(rate_)) {}
  void reset() {
    sub_.reset();
    stream_ = BlockStream();
  }
  void process(const float *mpx, size_t numSamples,
               const std::function<void(const RDSGroup &)> &onGroup) {
    if (!mpx || numSamples == 0) {
      return;
    }
    size_t offset = 0;
    while (offset < numSamples) {
      const size_t chunk = std::min<size_t>(8192, numSamples - offset);
      bits_.clear();
      sub_.chunkToBits(mpx + offset, chunk, bits_);
      all_bits_.insert(all_bits_.end(), bits_.begin(), bits_.end());
      for (uint8_t b : bits_) {
        stream_.pushBit(b != 0);
        if (!stream_.hasGroupReady()) {
          continue;
        }
        const RDSGroup g = packGroup(stream_.popGroup());
        if (onGroup) {
          onGroup(g);
        }
      }
      offset += chunk;
    }
  }
  // test hook: every demodulated bit since construction (not part of the reference API)
  std::vector<uint8_t> &allBits() { return all_bits_; }
  const SubcarrierSet &subcarriers() const { return sub_; }

private:
  int rate_;
  SubcarrierSet sub_;
  BlockStream stream_;
  std::vector<uint8_t> bits_;
  std::vector<uint8_t> all_bits_;
};

}  // namespace orc

#endif  // ORACLE_PIPELINE_HPP_
