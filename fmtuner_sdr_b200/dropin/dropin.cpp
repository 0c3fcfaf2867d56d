// dropin.cpp — the reference's block-processing classes re-implemented as thin wrappers over
// the engine's C ABI (include/fmgpu.h). Each object owns a one-channel engine sized for the
// largest dsp_block_samples the reference accepts (32768, src/main.cpp:680-681); the reference
// constructs exactly one of each (src/main.cpp:640-674,895). Creation failures surface as
// std::runtime_error, as the reference's wrappers do (src/dsp/liquid_primitives.cpp:30-32).
#include <algorithm>
#include <cmath>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "af_post_processor.h"
#include "dsp/liquid_primitives.h"
#include "dsp/runtime.h"
#include "fm_demod.h"
#include "fmgpu.h"
#include "rds_decoder.h"
#include "stereo_decoder.h"

namespace {

constexpr int kMaxBlock = 32768;

fmgpu_engine *makeEngine(int iqRate, int decimation, int outputRate, int tapsPerPhase = 0,
                         int attenDb = 0) {
  fmgpu_config cfg{};
  cfg.iq_rate = std::max(1, iqRate);
  cfg.decimation = std::max(1, decimation);
  cfg.output_rate = std::max(1, outputRate);
  cfg.block_samples = kMaxBlock;
  cfg.max_blocks = 1;
  cfg.w0_bandwidth_hz = 194000;  // FMDemod constructor default (fm_demod.cpp:33)
  cfg.bandwidth_hz = 309000;     // table index 0 == initial mode: keeps the constructor filter
  cfg.dsp_agc = 0;
  cfg.stereo_blend = 1;
  cfg.deemphasis = 1;            // constructors default to 75 us (fm_demod.cpp:46)
  cfg.stereo = 1;
  cfg.force_mono = 0;
  cfg.decim_taps_per_phase = tapsPerPhase;
  cfg.decim_atten_db = attenDb;
  fmgpu_engine *e = nullptr;
  const int rc = fmgpu_engine_create(&cfg, 1, 0, &e);
  if (rc != FMGPU_OK || !e) {
    throw std::runtime_error(std::string("failed to create fmgpu engine: ") + fmgpu_last_error(nullptr));
  }
  return e;
}

}  // namespace

// ---- FMDemod ---------------------------------------------------------------------------
FMDemod::FMDemod(int inputRate, int outputRate) : engine_(makeEngine(inputRate, 1, outputRate)) {}
FMDemod::~FMDemod() { fmgpu_engine_destroy(engine_); }

void FMDemod::process(const uint8_t *iq, float *audio, size_t numSamples) {
  fmgpu_demod_u8(engine_, 0, iq, nullptr, audio, numSamples);
}
void FMDemod::processComplex(const std::complex<float> *iq, float *audio, size_t numSamples) {
  fmgpu_demod_cf32(engine_, 0, reinterpret_cast<const float *>(iq), nullptr, audio, numSamples);
}
void FMDemod::processNoDownsample(const uint8_t *iq, float *audio, size_t numSamples) {
  fmgpu_demod_u8(engine_, 0, iq, audio, nullptr, numSamples);
}
size_t FMDemod::processSplit(const uint8_t *iq, float *mpxOut, float *monoOut, size_t numSamples) {
  return fmgpu_demod_u8(engine_, 0, iq, mpxOut, monoOut, numSamples);
}
size_t FMDemod::processSplitComplex(const std::complex<float> *iq, float *mpxOut, float *monoOut,
                                    size_t numSamples) {
  return fmgpu_demod_cf32(engine_, 0, reinterpret_cast<const float *>(iq), mpxOut, monoOut,
                          numSamples);
}
size_t FMDemod::downsampleAudio(const float *demod, float *audio, size_t numSamples) {
  return fmgpu_downsample_mono(engine_, 0, demod, audio, numSamples);
}
void FMDemod::reset() { fmgpu_reset(engine_, 0, FMGPU_RESET_DEMOD); }
void FMDemod::setDeemphasis(int tau_us) { fmgpu_set_deemphasis_us(engine_, 0, tau_us); }
void FMDemod::setDeviation(double deviation) { fmgpu_set_deviation_hz(engine_, deviation); }
void FMDemod::setBandwidthMode(int mode) { fmgpu_set_bandwidth_mode(engine_, 0, mode); }
void FMDemod::setBandwidthHz(int bwHz) { fmgpu_set_bandwidth_hz(engine_, 0, bwHz); }
void FMDemod::setW0BandwidthHz(int bwHz) { fmgpu_set_w0_bandwidth_hz(engine_, 0, bwHz); }
void FMDemod::setDspAgcMode(DspAgcMode mode) { fmgpu_set_agc_mode(engine_, 0, static_cast<int>(mode)); }
bool FMDemod::isClipping() const { return fmgpu_is_clipping(engine_, 0) != 0; }
float FMDemod::getClippingRatio() const { return fmgpu_clip_ratio(engine_, 0); }

// ---- StereoDecoder -----------------------------------------------------------------------
StereoDecoder::StereoDecoder(int inputRate, int outputRate)
    : engine_(makeEngine(inputRate, 1, outputRate > 0 ? outputRate : 32000)) {}
StereoDecoder::~StereoDecoder() { fmgpu_engine_destroy(engine_); }
size_t StereoDecoder::processAudio(const float *mono, float *left, float *right, size_t numSamples) {
  return fmgpu_stereo(engine_, 0, mono, left, right, numSamples);
}
void StereoDecoder::reset() { fmgpu_reset(engine_, 0, FMGPU_RESET_STEREO); }
void StereoDecoder::setForceStereo(bool force) { fmgpu_set_force_stereo(engine_, 0, force ? 1 : 0); }
void StereoDecoder::setForceMono(bool force) { fmgpu_set_force_mono(engine_, 0, force ? 1 : 0); }
void StereoDecoder::setBlendMode(BlendMode mode) { fmgpu_set_blend_mode(engine_, 0, static_cast<int>(mode)); }
int StereoDecoder::getPilotLevelTenthsKHz() const { return fmgpu_pilot_tenths(engine_, 0); }
bool StereoDecoder::isStereo() const { return fmgpu_is_stereo(engine_, 0) != 0; }

// ---- AFPostProcessor -----------------------------------------------------------------------
AFPostProcessor::AFPostProcessor(int inputRate, int outputRate)
    : engine_(makeEngine(inputRate, 1, outputRate)) {}
AFPostProcessor::~AFPostProcessor() { fmgpu_engine_destroy(engine_); }
void AFPostProcessor::reset() { fmgpu_reset(engine_, 0, FMGPU_RESET_AFPOST); }
void AFPostProcessor::setDeemphasis(int tau_us) { fmgpu_set_deemphasis_us(engine_, 0, tau_us); }
size_t AFPostProcessor::process(const float *inLeft, const float *inRight, size_t inSamples,
                                float *outLeft, float *outRight, size_t outCapacity) {
  return fmgpu_afpost(engine_, 0, inLeft, inRight, inSamples, outLeft, outRight, outCapacity);
}

// ---- RDSDecoder ------------------------------------------------------------------------------
RDSDecoder::RDSDecoder(int inputRate)
    : engine_(makeEngine(std::max(1, inputRate), 1, 32000)), rate_(std::max(1, inputRate)) {}
RDSDecoder::~RDSDecoder() { fmgpu_engine_destroy(engine_); }
void RDSDecoder::reset() { fmgpu_reset(engine_, 0, FMGPU_RESET_RDS); }
void RDSDecoder::process(const float *mpx, size_t numSamples,
                         const std::function<void(const RDSGroup &)> &onGroup) {
  if (!mpx || numSamples == 0) {
    return;
  }
  size_t offset = 0;
  std::vector<fmgpu_rds_group> groups(64);
  while (offset < numSamples) {
    const size_t chunk = std::min<size_t>(kMaxBlock, numSamples - offset);
    const size_t n = fmgpu_rds(engine_, 0, mpx + offset, chunk, groups.data(), groups.size());
    for (size_t i = 0; i < std::min(n, groups.size()); i++) {
      if (onGroup) {
        onGroup(RDSGroup{groups[i].a, groups[i].b, groups[i].c, groups[i].d, groups[i].errors});
      }
    }
    offset += chunk;
  }
}

// ---- ComplexDecimator ---------------------------------------------------------------------------
namespace fm_tuner::dsp::liquid {

ComplexDecimator::~ComplexDecimator() {
  if (engine_) {
    fmgpu_engine_destroy(engine_);
  }
}

void ComplexDecimator::init(std::uint32_t factor, std::uint32_t tapsPerPhase, float stopBandAtten) {
  if (factor == 0) {
    throw std::runtime_error("complex decimator factor must be >= 1");
  }
  if (engine_) {
    fmgpu_engine_destroy(engine_);
    engine_ = nullptr;
  }
  factor_ = factor;
  // factor 1 is a pure convert (liquid_primitives.cpp:468-478): it runs on the device as well
  engine_ = makeEngine(256000 * static_cast<int>(factor), static_cast<int>(factor), 32000,
                       static_cast<int>(std::max<std::uint32_t>(4, tapsPerPhase)),
                       static_cast<int>(std::lround(stopBandAtten)));
}

void ComplexDecimator::reset() {
  if (engine_) {
    fmgpu_reset(engine_, 0, FMGPU_RESET_DECIM);
  }
}

std::size_t ComplexDecimator::executeComplex(const uint8_t *iqIn, std::size_t inSamples,
                                             std::complex<float> *iqOut,
                                             std::size_t outCapacity) const {
  if (!iqIn || !iqOut || inSamples == 0 || outCapacity == 0) {
    return 0;
  }
  if (!engine_ && factor_ == 1) {
    engine_ = makeEngine(256000, 1, 32000);
  }
  if (!engine_) {
    return 0;
  }
  // (blocks longer than the engine was sized for go through in pieces)
  std::size_t done = 0;
  const std::size_t want = std::min(inSamples / factor_, outCapacity);
  while (done < want) {
    const std::size_t n = std::min<std::size_t>(kMaxBlock, want - done);
    const std::size_t got = fmgpu_decimate(engine_, 0, iqIn + 2 * done * factor_, n * factor_,
                                           reinterpret_cast<float *>(iqOut + done), n);
    if (got == 0) {
      break;
    }
    done += got;
  }
  return done;
}

std::size_t ComplexDecimator::execute(const uint8_t *iqIn, std::size_t inSamples, uint8_t *iqOut,
                                      std::size_t outCapacity) const {
  if (!iqIn || !iqOut || inSamples == 0 || outCapacity == 0) {
    return 0;
  }
  if (factor_ == 1) {
    const std::size_t n = std::min(inSamples, outCapacity);
    std::copy_n(iqIn, n * 2, iqOut);   // a byte copy, as in the reference (:431-435)
    return n;
  }
  if (!engine_) {
    return 0;
  }
  std::size_t done = 0;
  const std::size_t want = std::min(inSamples / factor_, outCapacity);
  while (done < want) {   // decimated and re-quantised on the device (fmgpu_decimate_u8)
    const std::size_t n = std::min<std::size_t>(kMaxBlock, want - done);
    const std::size_t got =
        fmgpu_decimate_u8(engine_, 0, iqIn + 2 * done * factor_, n * factor_, iqOut + 2 * done, n);
    if (got == 0) {
      break;
    }
    done += got;
  }
  return done;
}

}  // namespace fm_tuner::dsp::liquid

// ---- Runtime --------------------------------------------------------------------------------------
namespace fm_tuner::dsp {

const char *resetReasonName(ResetReason reason) {
  switch (reason) {
  case ResetReason::Start: return "start";
  case ResetReason::Stop: return "stop";
  case ResetReason::Retune: return "retune";
  case ResetReason::ScanRestore: return "scan_restore";
  }
  return "retune";
}

Runtime::Runtime(std::size_t blockSize, bool verbose)
    : blockSize_(std::max<std::size_t>(1, blockSize)), verbose_(verbose) {
  // self-test: the engine must be able to create and destroy a channel on device 0
  try {
    fmgpu_engine_destroy(makeEngine(256000, 1, 32000));
  } catch (const std::exception &ex) {
    throw std::runtime_error(std::string("failed to initialize fmgpu backend: ") + ex.what());
  }
  if (verbose_) {
    std::cout << "[DSP] fmgpu (CUDA) primitives initialized\n";
  }
}

Runtime::~Runtime() = default;

void Runtime::addResetHandler(std::function<void()> handler) {
  if (handler) {
    handlers_.push_back(std::move(handler));
  }
}

void Runtime::reset(ResetReason reason) const {
  if (verbose_) {
    std::cout << "[DSP] reset reason=" << resetReasonName(reason) << "\n";
  }
  for (const auto &h : handlers_) {
    h();
  }
}

}  // namespace fm_tuner::dsp
