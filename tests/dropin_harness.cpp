// tests/dropin_harness.cpp — drives the drop-in classes exactly the way the reference's main
// loop does (src/main.cpp:640-710 construction, :1232-1308 per-block body, :894-927 RDS) and
// dumps the results for tests/test_gpu_dropin.py to compare with the CPU oracle.
//   usage: dropin_harness <iq.u8> <iq_rate> <decimation> <block> <out_prefix>
#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "af_post_processor.h"
#include "dsp/liquid_primitives.h"
#include "dsp/runtime.h"
#include "fm_demod.h"
#include "rds_decoder.h"
#include "stereo_decoder.h"

int main(int argc, char **argv) {
  if (argc < 6) {
    return 2;
  }
  const int iqRate = std::atoi(argv[2]);
  const int decim = std::atoi(argv[3]);
  const size_t BUF = static_cast<size_t>(std::atoi(argv[4]));
  const std::string prefix = argv[5];
  const int INPUT_RATE = iqRate / decim, OUTPUT_RATE = 32000;

  FILE *f = std::fopen(argv[1], "rb");
  if (!f) {
    return 3;
  }
  std::vector<uint8_t> iq;
  uint8_t tmp[65536];
  size_t got;
  while ((got = std::fread(tmp, 1, sizeof(tmp), f)) > 0) {
    iq.insert(iq.end(), tmp, tmp + got);
  }
  std::fclose(f);

  FMDemod demod(INPUT_RATE, OUTPUT_RATE);
  demod.setW0BandwidthHz(194000);
  demod.setDspAgcMode(FMDemod::DspAgcMode::Off);
  StereoDecoder stereo(INPUT_RATE, OUTPUT_RATE);
  stereo.setBlendMode(StereoDecoder::BlendMode::Normal);
  AFPostProcessor afPost(INPUT_RATE, OUTPUT_RATE);
  fm_tuner::dsp::liquid::ComplexDecimator iqDecimator;
  const uint32_t f32 = static_cast<uint32_t>(decim);
  iqDecimator.init(f32, (f32 >= 8U) ? 28U : ((f32 >= 4U) ? 20U : 12U), 80.0f);
  fm_tuner::dsp::Runtime dspRuntime(BUF, false);
  dspRuntime.addResetHandler([&]() {
    demod.reset();
    stereo.reset();
    afPost.reset();
    iqDecimator.reset();
  });
  afPost.setDeemphasis(50);
  demod.setDeemphasis(50);
  stereo.setForceMono(false);
  demod.setBandwidthHz(0);
  RDSDecoder rds(INPUT_RATE);
  dspRuntime.reset(fm_tuner::dsp::ResetReason::Start);

  std::vector<std::complex<float>> cplx(BUF);
  std::vector<float> mpx(BUF), sl(BUF), sr(BUF), al(BUF), ar(BUF);
  FILE *fa = std::fopen((prefix + ".audio.f32").c_str(), "wb");
  FILE *fs = std::fopen((prefix + ".status.txt").c_str(), "w");
  FILE *fg = std::fopen((prefix + ".groups.txt").c_str(), "w");
  const size_t perBlock = BUF * decim * 2;
  const size_t nblk = iq.size() / perBlock;
  for (size_t b = 0; b < nblk; b++) {
    const uint8_t *in = iq.data() + b * perBlock;
    size_t n = BUF;
    if (decim > 1) {
      n = iqDecimator.executeComplex(in, BUF * decim, cplx.data(), BUF);
      demod.processSplitComplex(cplx.data(), mpx.data(), nullptr, n);
    } else {
      demod.processSplit(in, mpx.data(), nullptr, n);
    }
    int ng = 0;
    rds.process(mpx.data(), n, [&](const RDSGroup &g) {
      std::fprintf(fg, "%zu %u %u %u %u %u\n", b, g.blockA, g.blockB, g.blockC, g.blockD, g.errors);
      ng++;
    });
    const size_t ss = stereo.processAudio(mpx.data(), sl.data(), sr.data(), n);
    const size_t out = afPost.process(sl.data(), sr.data(), ss, al.data(), ar.data(), BUF);
    for (size_t i = 0; i < out; i++) {
      al[i] = std::clamp(al[i], -1.0f, 1.0f);
      ar[i] = std::clamp(ar[i], -1.0f, 1.0f);
    }
    std::fwrite(al.data(), sizeof(float), out, fa);
    std::fwrite(ar.data(), sizeof(float), out, fa);
    std::fprintf(fs, "%zu %d %d %.9g %d\n", out, stereo.isStereo() ? 1 : 0,
                 stereo.getPilotLevelTenthsKHz(), demod.getClippingRatio(), ng);
  }
  std::fclose(fa);
  std::fclose(fs);
  std::fclose(fg);

  // ---- the rest of the public surface, on fresh objects ---------------------------------
  {
    FILE *fx = std::fopen((prefix + ".extra.bin").c_str(), "wb");
    const size_t n = std::min<size_t>(4096, iq.size() / 2 / std::max(1, decim));
    // ComplexDecimator::execute (re-quantised uint8 output)
    if (decim > 1) {
      fm_tuner::dsp::liquid::ComplexDecimator d2;
      d2.init(f32, (f32 >= 8U) ? 28U : ((f32 >= 4U) ? 20U : 12U), 80.0f);
      std::vector<uint8_t> q(2 * n);
      const size_t k = d2.execute(iq.data(), n * decim, q.data(), n);
      std::fwrite(q.data(), 1, 2 * k, fx);
    }
    // FMDemod::process (mono audio), processNoDownsample, downsampleAudio, setDeviation,
    // setBandwidthMode on raw uint8 IQ at the DSP rate
    FMDemod d3(INPUT_RATE, OUTPUT_RATE);
    d3.setBandwidthMode(9);
    d3.setDeviation(50000.0);
    std::vector<float> a(n), b(n), c(n);
    d3.process(iq.data(), a.data(), n);
    std::fwrite(a.data(), sizeof(float), n * OUTPUT_RATE / INPUT_RATE - 2, fx);
    d3.processNoDownsample(iq.data() + 2 * n, b.data(), n);
    std::fwrite(b.data(), sizeof(float), n, fx);
    const size_t k2 = d3.downsampleAudio(b.data(), c.data(), n);
    std::fwrite(c.data(), sizeof(float), k2, fx);
    // a default-constructed ComplexDecimator (factor 1, never initialised): the pure convert of
    // liquid_primitives.cpp:468-478
    fm_tuner::dsp::liquid::ComplexDecimator d4;
    std::vector<std::complex<float>> conv(n);
    const size_t k3 = d4.executeComplex(iq.data(), n, conv.data(), n);
    std::fwrite(conv.data(), sizeof(std::complex<float>), k3, fx);
    std::fclose(fx);
  }
  return 0;
}
