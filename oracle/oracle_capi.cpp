// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" surface over the CPU oracle so tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline leg can drive it through ctypes. Built twice by
// oracle/Makefile: liboracle_libm.so (libm transcendentals, faithful to the
// reference call sites) and liboracle_fm.so (-DORACLE_FM_MATH, transcendental
// kernels shared with the engine, for bit-for-bit comparison).
//
// The Channel class restates the reference's per-block glue, src/main.cpp:1232-1308
// (decimate -> FMDemod -> RDS (synchronously, Appendix B.9) -> StereoDecoder ->
// AFPostProcessor -> clamp), with INPUT_RATE generalised from the compile-time
// 256000 (src/main.cpp:72) to iq_rate / decimation.
#include <cstdint>
#include <cstring>
#include <memory>
#include <new>
#include <vector>

#ifdef ORACLE_USE_REFERENCE
// Third build (oracle/Makefile, target ref): the SAME harness over the reference's own classes,
// compiled unmodified from /root/reference over liquid_shim/ -> oracle/_ref/libfmref.so.
#include <algorithm>
#include <complex>

#include "af_post_processor.h"
#include "dsp/liquid_primitives.h"
#include "fm_demod.h"
#include "rds_decoder.h"
#include "redsea_port/dsp/subcarrier.hh"
#include "stereo_decoder.h"

namespace orc {
using cf32 = std::complex<float>;
using ::AFPostProcessor;
using ::FMDemod;
using ::RDSDecoder;
using ::RDSGroup;
using ::StereoDecoder;
using ComplexDecimator = fm_tuner::dsp::liquid::ComplexDecimator;
namespace m {
inline const char *name() { return "reference"; }
}  // namespace m
}  // namespace orc
#else
#include "pipeline.hpp"
#endif

extern "C" {

struct orc_config {
  int32_t iq_rate;        // 256000|1024000|2048000 (reference) or 240000|2400000 (class level)
  int32_t decimation;     // 1, 4, 8 (reference) or 10
  int32_t block_samples;  // dsp_block_samples at the DSP rate (default 8192)
  int32_t w0_bandwidth_hz;  // processing.w0_bandwidth_hz (default 194000)
  int32_t bandwidth_hz;     // XDR 'W' value applied at start (0 => W0)
  int32_t dsp_agc;          // 0 off, 1 fast, 2 slow
  int32_t stereo_blend;     // 0 soft, 1 normal, 2 aggressive
  int32_t deemphasis;       // 0 = 50us, 1 = 75us, 2 = off   (include/config.h:37)
  int32_t stereo;           // processing.stereo
  int32_t force_mono;
};

struct orc_block_status {
  int32_t n_audio;       // 32 kHz frames produced by this block
  int32_t stereo;        // StereoDecoder::isStereo() after the block
  int32_t pilot_tenths;  // getPilotLevelTenthsKHz()
  float clip_ratio;      // FMDemod::getClippingRatio()
  int32_t n_groups;      // RDS groups emitted during this block
};

struct orc_group {
  uint16_t a, b, c, d;
  uint8_t errors;
  uint8_t pad[3];
  uint32_t block_index;
};

}  // extern "C"

namespace {

using namespace orc;

class Channel {
public:
  explicit Channel(const orc_config &c)
      : cfg_(c), fs_(c.iq_rate / std::max(1, c.decimation)), demod_(fs_, 32000),
        stereo_(fs_, 32000), afpost_(fs_, 32000), rds_(fs_) {
    demod_.setW0BandwidthHz(c.w0_bandwidth_hz);  // main.cpp:641
    demod_.setDspAgcMode(static_cast<FMDemod::DspAgcMode>(c.dsp_agc));
    stereo_.setBlendMode(static_cast<StereoDecoder::BlendMode>(c.stereo_blend));
    const uint32_t f = static_cast<uint32_t>(c.decimation);
    const uint32_t tpp = (f >= 8U) ? 28U : ((f >= 4U) ? 20U : 12U);  // main.cpp:672-674
    decim_.init(f, tpp, 80.0f);
    const int us = (c.deemphasis == 0) ? 50 : ((c.deemphasis == 1) ? 75 : 0);  // main.cpp:699-708
    afpost_.setDeemphasis(us);
    demod_.setDeemphasis(us);
    stereo_.setForceMono(c.force_mono != 0);
    demod_.setBandwidthHz(c.bandwidth_hz);  // main.cpp:710
    const size_t n = static_cast<size_t>(c.block_samples);
    cplx_.resize(n);
    mpx_.resize(n);
    sl_.resize(n);
    sr_.resize(n);
  }

  void reset(bool dsp, bool rds) {
    if (dsp) {  // main.cpp:686-691
      demod_.reset();
      stereo_.reset();
      afpost_.reset();
      decim_.reset();
    }
    if (rds) {
      rds_.reset();
#ifdef ORACLE_USE_REFERENCE
      if (tap_) {
        tap_->reset();  // RDSDecoder::Impl::reset, rds_decoder.cpp:23-27
      }
#endif
    }
  }

#ifdef ORACLE_USE_REFERENCE
  // RDSDecoder keeps its SubcarrierSet private and hands out groups only. For bit-level
  // comparisons a second, public redsea::SubcarrierSet is fed the same MPX in the same chunks
  // (rds_decoder.cpp:74-93); off unless orc_channel_enable_bits_tap was called.
  void enableTap() {
    if (!tap_) {
      tap_ = std::make_unique<redsea::SubcarrierSet>(static_cast<float>(std::max(1, fs_)));
    }
  }
  void tapBits(const float *mpx, size_t n) {
    if (!tap_ || !mpx) {
      return;
    }
    size_t offset = 0;
    while (offset < n) {
      const size_t chunk = std::min(static_cast<size_t>(redsea::kInputChunkSize), n - offset);
      auto input = std::make_unique<redsea::MPXBuffer>();
      input->used_size = chunk;
      std::memcpy(input->data.data(), mpx + offset, chunk * sizeof(float));
      const redsea::BitBuffer bits = tap_->chunkToBits(*input, 1);
      for (const redsea::TimedBit &b : bits.bits[0]) {
        all_bits_.push_back(b.value ? 1 : 0);
      }
      offset += chunk;
    }
  }
  std::unique_ptr<redsea::SubcarrierSet> tap_;
  std::vector<uint8_t> all_bits_;
  std::vector<uint8_t> &allBits() { return all_bits_; }
#else
  std::vector<uint8_t> &allBits() { return rds_.allBits(); }
#endif

  // one logical block: iq holds block_samples * decimation IQ pairs
  size_t processBlock(const uint8_t *iq, float *outL, float *outR, size_t cap,
                      orc_block_status *st, std::vector<orc_group> &groups, uint32_t block_index,
                      float *dbg_dec, float *dbg_mpx, float *dbg_sl, float *dbg_sr) {
    const size_t n = static_cast<size_t>(cfg_.block_samples);
    size_t demodSamples = n;
    const bool decimate = cfg_.decimation > 1;
    if (decimate) {
      demodSamples = decim_.executeComplex(iq, n * cfg_.decimation, cplx_.data(), n);
      if (dbg_dec) {
        std::memcpy(dbg_dec, cplx_.data(), demodSamples * sizeof(cf32));
      }
    }
    size_t outSamples = 0;
    const size_t g0 = groups.size();
    auto onGroup = [&](const RDSGroup &g) {
      groups.push_back(orc_group{g.blockA, g.blockB, g.blockC, g.blockD, g.errors, {0, 0, 0},
                                 block_index});
    };
    bool stereoDetected = false;
    int pilotTenths = 0;
    if (!cfg_.stereo) {  // main.cpp:1266-1279
      std::vector<float> mono(cap);
      outSamples = decimate ? demod_.processSplitComplex(cplx_.data(), mpx_.data(), mono.data(),
                                                         demodSamples)
                            : demod_.processSplit(iq, mpx_.data(), mono.data(), demodSamples);
      rds_.process(mpx_.data(), demodSamples, onGroup);
#ifdef ORACLE_USE_REFERENCE
      tapBits(mpx_.data(), demodSamples);
#endif
      for (size_t i = 0; i < outSamples; i++) {
        const float v = mono[i] * 0.5f;
        outL[i] = v;
        outR[i] = v;
      }
    } else {  // main.cpp:1280-1297
      if (decimate) {
        demod_.processSplitComplex(cplx_.data(), mpx_.data(), nullptr, demodSamples);
      } else {
        demod_.processSplit(iq, mpx_.data(), nullptr, demodSamples);
      }
      rds_.process(mpx_.data(), demodSamples, onGroup);
#ifdef ORACLE_USE_REFERENCE
      tapBits(mpx_.data(), demodSamples);
#endif
      const size_t ss = stereo_.processAudio(mpx_.data(), sl_.data(), sr_.data(), demodSamples);
      outSamples = afpost_.process(sl_.data(), sr_.data(), ss, outL, outR, std::min(cap, n));
      stereoDetected = stereo_.isStereo();
      pilotTenths = stereo_.getPilotLevelTenthsKHz();
      if (dbg_sl) {
        std::memcpy(dbg_sl, sl_.data(), ss * sizeof(float));
      }
      if (dbg_sr) {
        std::memcpy(dbg_sr, sr_.data(), ss * sizeof(float));
      }
    }
    if (dbg_mpx) {
      std::memcpy(dbg_mpx, mpx_.data(), demodSamples * sizeof(float));
    }
    for (size_t i = 0; i < outSamples; i++) {  // main.cpp:1305-1308
      outL[i] = std::clamp(outL[i], -1.0f, 1.0f);
      outR[i] = std::clamp(outR[i], -1.0f, 1.0f);
    }
    if (st) {
      st->n_audio = static_cast<int32_t>(outSamples);
      st->stereo = stereoDetected ? 1 : 0;
      st->pilot_tenths = pilotTenths;
      st->clip_ratio = demod_.getClippingRatio();
      st->n_groups = static_cast<int32_t>(groups.size() - g0);
    }
    return outSamples;
  }

  orc_config cfg_;
  int fs_;
  FMDemod demod_;
  StereoDecoder stereo_;
  AFPostProcessor afpost_;
  ComplexDecimator decim_;
  RDSDecoder rds_;
  std::vector<cf32> cplx_;
  std::vector<float> mpx_, sl_, sr_;
};

template <typename F> int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &) {
    return -1;
  }
}

size_t copyOut(const std::vector<float> &v, float *out, size_t cap) {
  if (out) {
    std::memcpy(out, v.data(), std::min(cap, v.size()) * sizeof(float));
  }
  return v.size();
}

}  // namespace

extern "C" {

const char *orc_math_name() { return orc::m::name(); }

// ---- whole-channel harness ---------------------------------------------------
void *orc_channel_create(const orc_config *cfg) {
  Channel *c = nullptr;
  if (guarded([&] { c = new Channel(*cfg); }) != 0) {
    return nullptr;
  }
  return c;
}
void orc_channel_destroy(void *h) { delete static_cast<Channel *>(h); }
void orc_channel_reset(void *h, int dsp, int rds) { static_cast<Channel *>(h)->reset(dsp, rds); }
void orc_channel_set_bandwidth_hz(void *h, int bw) {
  static_cast<Channel *>(h)->demod_.setBandwidthHz(bw);
}
void orc_channel_set_force_mono(void *h, int f) {
  static_cast<Channel *>(h)->stereo_.setForceMono(f != 0);
}
void orc_channel_set_force_stereo(void *h, int f) {
  static_cast<Channel *>(h)->stereo_.setForceStereo(f != 0);
}
void orc_channel_set_deemphasis(void *h, int mode) {
  Channel *c = static_cast<Channel *>(h);
  const int us = (mode == 0) ? 50 : ((mode == 1) ? 75 : 0);
  c->afpost_.setDeemphasis(us);
  c->demod_.setDeemphasis(us);
}

// Processes n_blocks logical blocks. iq: n_blocks*block*decim IQ pairs. Audio is
// appended block after block into outL/outR (capacity out_cap frames in total).
// Debug taps (may be null): dec [n_blocks*block] cf32, mpx/sl/sr [n_blocks*block] f32.
// Returns total audio frames, or -1 if a capacity was exceeded.
long orc_channel_process(void *h, const uint8_t *iq, size_t n_blocks, float *outL, float *outR,
                         size_t out_cap, orc_block_status *status, orc_group *groups,
                         size_t group_cap, size_t *n_groups, float *dbg_dec, float *dbg_mpx,
                         float *dbg_sl, float *dbg_sr) {
  Channel *c = static_cast<Channel *>(h);
  const size_t n = static_cast<size_t>(c->cfg_.block_samples);
  const size_t iq_per_block = n * static_cast<size_t>(c->cfg_.decimation) * 2;
  std::vector<orc_group> gv;
  size_t total = 0;
  for (size_t b = 0; b < n_blocks; b++) {
    if (total + n > out_cap) {
      return -1;
    }
    total += c->processBlock(iq + b * iq_per_block, outL + total, outR + total, out_cap - total,
                             status ? &status[b] : nullptr, gv, static_cast<uint32_t>(b),
                             dbg_dec ? dbg_dec + 2 * b * n : nullptr,
                             dbg_mpx ? dbg_mpx + b * n : nullptr, dbg_sl ? dbg_sl + b * n : nullptr,
                             dbg_sr ? dbg_sr + b * n : nullptr);
  }
  if (n_groups) {
    *n_groups = gv.size();
  }
  if (groups) {
    if (gv.size() > group_cap) {
      return -1;
    }
    std::memcpy(groups, gv.data(), gv.size() * sizeof(orc_group));
  }
  return static_cast<long>(total);
}

// reference build only: start collecting demodulated bits (no-op for the restated oracle, which
// always collects them)
void orc_channel_enable_bits_tap(void *h) {
#ifdef ORACLE_USE_REFERENCE
  static_cast<Channel *>(h)->enableTap();
#else
  (void)h;
#endif
}

// every RDS bit demodulated so far (before block sync)
size_t orc_channel_rds_bits(void *h, uint8_t *out, size_t cap) {
  auto &v = static_cast<Channel *>(h)->allBits();
  if (out) {
    std::memcpy(out, v.data(), std::min(cap, v.size()));
  }
  return v.size();
}

// ---- class-level handles (mirror the reference's public methods) -----------------
void *orc_decim_create(uint32_t factor, uint32_t tpp, float atten) {
  ComplexDecimator *d = nullptr;
  if (guarded([&] {
        d = new ComplexDecimator();
        d->init(factor, tpp, atten);
      }) != 0) {
    return nullptr;
  }
  return d;
}
void orc_decim_destroy(void *h) { delete static_cast<ComplexDecimator *>(h); }
void orc_decim_reset(void *h) { static_cast<ComplexDecimator *>(h)->reset(); }
size_t orc_decim_execute_complex(void *h, const uint8_t *iq, size_t n_in, float *out_cf32,
                                 size_t cap) {
  return static_cast<ComplexDecimator *>(h)->executeComplex(iq, n_in,
                                                            reinterpret_cast<cf32 *>(out_cf32), cap);
}
size_t orc_decim_execute_u8(void *h, const uint8_t *iq, size_t n_in, uint8_t *out, size_t cap) {
  return static_cast<ComplexDecimator *>(h)->execute(iq, n_in, out, cap);
}

void *orc_demod_create(int in_rate, int out_rate) {
  FMDemod *d = nullptr;
  if (guarded([&] { d = new FMDemod(in_rate, out_rate); }) != 0) {
    return nullptr;
  }
  return d;
}
void orc_demod_destroy(void *h) { delete static_cast<FMDemod *>(h); }
void orc_demod_reset(void *h) { static_cast<FMDemod *>(h)->reset(); }
void orc_demod_set_w0(void *h, int bw) { static_cast<FMDemod *>(h)->setW0BandwidthHz(bw); }
void orc_demod_set_bandwidth_hz(void *h, int bw) { static_cast<FMDemod *>(h)->setBandwidthHz(bw); }
void orc_demod_set_bandwidth_mode(void *h, int mode) {
  static_cast<FMDemod *>(h)->setBandwidthMode(mode);
}
void orc_demod_set_agc(void *h, int mode) {
  static_cast<FMDemod *>(h)->setDspAgcMode(static_cast<FMDemod::DspAgcMode>(mode));
}
void orc_demod_set_deemphasis(void *h, int us) { static_cast<FMDemod *>(h)->setDeemphasis(us); }
size_t orc_demod_process_split(void *h, const uint8_t *iq, float *mpx, float *mono, size_t n) {
  return static_cast<FMDemod *>(h)->processSplit(iq, mpx, mono, n);
}
size_t orc_demod_process_split_complex(void *h, const float *iq_cf32, float *mpx, float *mono,
                                       size_t n) {
  return static_cast<FMDemod *>(h)->processSplitComplex(reinterpret_cast<const cf32 *>(iq_cf32),
                                                        mpx, mono, n);
}
void orc_demod_set_deviation(void *h, double dev) { static_cast<FMDemod *>(h)->setDeviation(dev); }
size_t orc_demod_downsample(void *h, const float *mpx, float *audio, size_t n) {
  return static_cast<FMDemod *>(h)->downsampleAudio(mpx, audio, n);
}
float orc_demod_clip_ratio(void *h) { return static_cast<FMDemod *>(h)->getClippingRatio(); }
int orc_demod_is_clipping(void *h) { return static_cast<FMDemod *>(h)->isClipping() ? 1 : 0; }

void *orc_stereo_create(int in_rate) {
  StereoDecoder *d = nullptr;
  if (guarded([&] { d = new StereoDecoder(in_rate, 32000); }) != 0) {
    return nullptr;
  }
  return d;
}
void orc_stereo_destroy(void *h) { delete static_cast<StereoDecoder *>(h); }
void orc_stereo_reset(void *h) { static_cast<StereoDecoder *>(h)->reset(); }
void orc_stereo_set_blend(void *h, int mode) {
  static_cast<StereoDecoder *>(h)->setBlendMode(static_cast<StereoDecoder::BlendMode>(mode));
}
void orc_stereo_set_force_mono(void *h, int f) { static_cast<StereoDecoder *>(h)->setForceMono(f); }
void orc_stereo_set_force_stereo(void *h, int f) {
  static_cast<StereoDecoder *>(h)->setForceStereo(f);
}
size_t orc_stereo_process(void *h, const float *mpx, float *l, float *r, size_t n) {
  return static_cast<StereoDecoder *>(h)->processAudio(mpx, l, r, n);
}
int orc_stereo_is_stereo(void *h) { return static_cast<StereoDecoder *>(h)->isStereo() ? 1 : 0; }
int orc_stereo_pilot_tenths(void *h) {
  return static_cast<StereoDecoder *>(h)->getPilotLevelTenthsKHz();
}

void *orc_afpost_create(int in_rate, int out_rate) {
  AFPostProcessor *d = nullptr;
  if (guarded([&] { d = new AFPostProcessor(in_rate, out_rate); }) != 0) {
    return nullptr;
  }
  return d;
}
void orc_afpost_destroy(void *h) { delete static_cast<AFPostProcessor *>(h); }
void orc_afpost_reset(void *h) { static_cast<AFPostProcessor *>(h)->reset(); }
void orc_afpost_set_deemphasis(void *h, int us) {
  static_cast<AFPostProcessor *>(h)->setDeemphasis(us);
}
size_t orc_afpost_process(void *h, const float *l, const float *r, size_t n, float *ol, float *or_,
                          size_t cap) {
  return static_cast<AFPostProcessor *>(h)->process(l, r, n, ol, or_, cap);
}

void *orc_rds_create(int in_rate) {
  RDSDecoder *d = nullptr;
  if (guarded([&] { d = new RDSDecoder(in_rate); }) != 0) {
    return nullptr;
  }
  return d;
}
void orc_rds_destroy(void *h) { delete static_cast<RDSDecoder *>(h); }
void orc_rds_reset(void *h) { static_cast<RDSDecoder *>(h)->reset(); }
size_t orc_rds_process(void *h, const float *mpx, size_t n, orc_group *out, size_t cap) {
  size_t k = 0;
  static_cast<RDSDecoder *>(h)->process(mpx, n, [&](const RDSGroup &g) {
    if (k < cap) {
      out[k] = orc_group{g.blockA, g.blockB, g.blockC, g.blockD, g.errors, {0, 0, 0}, 0};
    }
    k++;
  });
  return k;
}
#ifndef ORACLE_USE_REFERENCE
size_t orc_rds_bits(void *h, uint8_t *out, size_t cap) {
  auto &v = static_cast<RDSDecoder *>(h)->allBits();
  if (out) {
    std::memcpy(out, v.data(), std::min(cap, v.size()));
  }
  return v.size();
}

// ---- integer block synchroniser on a raw bit stream ------------------------------
size_t orc_blockstream_run(const uint8_t *bits, size_t n_bits, orc_group *out, size_t cap) {
  BlockStream bs;
  size_t k = 0;
  for (size_t i = 0; i < n_bits; i++) {
    bs.pushBit(bits[i] != 0);
    if (bs.hasGroupReady()) {
      const RDSGroup g = packGroup(bs.popGroup());
      if (k < cap) {
        out[k] = orc_group{g.blockA, g.blockB, g.blockC, g.blockD, g.errors, {0, 0, 0},
                           static_cast<uint32_t>(i)};
      }
      k++;
    }
  }
  return k;
}
uint32_t orc_rds_syndrome(uint32_t v) { return rdsSyndrome(v); }

// ---- design getters (to compare with the engine's own design code) ---------------
// which: 0 decimator taps (a=factor, b=tpp, fa=atten)
//        1 channel filter taps for FMDemod(rate=a) after setW0(194000), setBandwidthHz(b)
//        2 pilot band-pass taps (rate=a)            3 audio low-pass taps (rate=a)
//        4 resampler bank [32][2m] (fa = ratio, a = m)  5 RDS low-pass taps
//        6 symsync MF bank [32][18]                 7 symsync dMF bank
size_t orc_design(int which, int a, int b, float fa, float *out, size_t cap, float *scale) {
  std::vector<float> v;
  float sc = 1.0f;
  const int rc = guarded([&] {
    switch (which) {
    case 0: {
      ComplexDecimator d;
      d.init(static_cast<uint32_t>(a), static_cast<uint32_t>(b), fa);
      v = d.taps();
      sc = d.scale();
      break;
    }
    case 1: {
      FMDemod d(a, 32000);
      d.setW0BandwidthHz(194000);
      d.setBandwidthHz(b);
      v = d.iqFilter().taps();
      sc = d.iqFilter().scale();
      break;
    }
    case 2: {
      StereoDecoder d(a, 32000);
      v = d.pilotFilter().taps();
      sc = d.pilotFilter().scale();
      break;
    }
    case 3: {
      StereoDecoder d(a, 32000);
      v = d.audioFilter().taps();
      sc = d.audioFilter().scale();
      break;
    }
    case 4: {
      Resamp r;
      r.create(1.0f, static_cast<unsigned>(a), 0.47f, 60.0f, 32);
      r.set_rate(fa);
      v = r.bank();
      sc = static_cast<float>(r.step());
      break;
    }
    case 5: {
      SubcarrierSet s(240000.f);
      v = s.lpf().taps();
      sc = s.lpf().scale();
      break;
    }
    case 6: {
      SubcarrierSet s(240000.f);
      v = s.symsync().mf_bank();
      sc = s.symsync().sos_b0();
      break;
    }
    case 7: {
      SubcarrierSet s(240000.f);
      v = s.symsync().dmf_bank();
      sc = s.symsync().sos_a1();
      break;
    }
    default: break;
    }
  });
  if (rc != 0) {
    return 0;
  }
  if (scale) {
    *scale = sc;
  }
  return copyOut(v, out, cap);
}

// ---- scalar math under test (the oracle's own dispatch) ----------------------------
void orc_math_sincos(const float *x, float *s, float *c, size_t n) {
  for (size_t i = 0; i < n; i++) {
    s[i] = orc::m::sin(x[i]);
    c[i] = orc::m::cos(x[i]);
  }
}
void orc_math_atan2(const float *y, const float *x, float *r, size_t n) {
  for (size_t i = 0; i < n; i++) {
    r[i] = orc::m::atan2(y[i], x[i]);
  }
}
void orc_math_exp(const float *x, float *r, size_t n) {
  for (size_t i = 0; i < n; i++) {
    r[i] = orc::m::exp(x[i]);
  }
}
void orc_math_log(const float *x, float *r, size_t n) {
  for (size_t i = 0; i < n; i++) {
    r[i] = orc::m::log(x[i]);
  }
}
uint32_t orc_nco_constrain(float x) { return orc::nco_constrain(x); }
#endif  // !ORACLE_USE_REFERENCE

}  // extern "C"
