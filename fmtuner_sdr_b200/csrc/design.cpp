// design.cpp — see design.h. Host only; compiled with -ffp-contract=off.
#include "design.h"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace fmdesign {

namespace {

// modified Bessel function of the first kind, order 0 (power series)
double i0(double x) {
  const double half = 0.5 * x;
  double term = 1.0;
  double acc = 1.0;
  for (int k = 1; k < 200; k++) {
    term *= half / static_cast<double>(k);
    const double sq = term * term;
    acc += sq;
    if (sq < 1e-22 * acc) {
      break;
    }
  }
  return acc;
}

double kaiserBeta(double As) {
  As = std::fabs(As);
  if (As > 50.0) {
    return 0.1102 * (As - 8.7);
  }
  if (As > 21.0) {
    return 0.5842 * std::pow(As - 21.0, 0.4) + 0.07886 * (As - 21.0);
  }
  return 0.0;
}

double sincPi(double x) {
  if (std::fabs(x) < 0.01) {
    return std::cos(M_PI * x / 2.0) * std::cos(M_PI * x / 4.0) * std::cos(M_PI * x / 8.0);
  }
  return std::sin(M_PI * x) / (M_PI * x);
}

}  // namespace

std::vector<float> kaiserLowpass(unsigned n, float fc, float As, float mu) {
  if (n == 0 || !(fc > 0.0f) || !(fc <= 0.5f)) {
    throw std::runtime_error("kaiserLowpass: invalid arguments");
  }
  const double beta = kaiserBeta(As);
  const double centre = static_cast<double>(n - 1) / 2.0;
  const double denom = i0(beta);
  std::vector<float> h(n);
  for (unsigned i = 0; i < n; i++) {
    const double t = static_cast<double>(i) - centre + mu;
    const double tw = static_cast<double>(i) - centre;
    const double r = 2.0 * tw / static_cast<double>(n - 1);
    const double w = i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / denom;
    h[i] = static_cast<float>(sincPi(2.0 * static_cast<double>(fc) * t) * w);
  }
  return h;
}

std::vector<float> shiftedBandpass(unsigned n, float fc, float As, float center) {
  std::vector<float> taps = kaiserLowpass(n, fc, As, 0.0f);
  const int mid = static_cast<int>(n / 2);
  constexpr float kTwoPi = 6.28318530717958647692f;
  for (unsigned i = 0; i < n; i++) {
    const float phase = kTwoPi * center * static_cast<float>(static_cast<int>(i) - mid);
    taps[i] = 2.0f * taps[i] * std::cos(phase);
  }
  double sumAbs = 0.0;
  for (float t : taps) {
    sumAbs += std::abs(t);
  }
  if (sumAbs > 1e-12) {
    const float inv = static_cast<float>(1.0 / sumAbs);
    for (float &t : taps) {
      t *= inv;
    }
  }
  return taps;
}

std::vector<float> rootRaisedCosine(unsigned k, unsigned m, float betaF) {
  const unsigned n = 2 * k * m + 1;
  const double beta = betaF;
  std::vector<float> h(n);
  for (unsigned i = 0; i < n; i++) {
    const double z = static_cast<double>(i) / static_cast<double>(k) - static_cast<double>(m);
    double v;
    if (std::fabs(z) < 1e-5) {
      v = 1.0 - beta + 4.0 * beta / M_PI;
    } else {
      double g = 1.0 - 16.0 * beta * beta * z * z;
      const double lead = 4.0 * beta / (M_PI * g);
      g *= g;
      if (g < 1e-5) {
        v = beta / std::sqrt(2.0) *
            ((1.0 + 2.0 / M_PI) * std::sin(0.25 * M_PI / beta) +
             (1.0 - 2.0 / M_PI) * std::cos(0.25 * M_PI / beta));
      } else {
        const double c = std::cos((1.0 + beta) * M_PI * z);
        const double s = std::sin((1.0 - beta) * M_PI * z);
        v = lead * (c + (s * (1.0 / (4.0 * beta * z))));
      }
    }
    h[i] = static_cast<float>(v);
  }
  return h;
}

uint32_t resamplerStep(float rate) {
  const float q = static_cast<float>(1 << 24) / rate;  // int / float: float division
  return static_cast<uint32_t>(std::round(static_cast<double>(q)));
}

ResamplerDesign resampler(float rate, unsigned m, float fc, float As, unsigned npfb) {
  if (!(rate > 0.0f) || m == 0 || !(fc > 0.0f) || !(fc < 0.5f) || npfb == 0) {
    throw std::runtime_error("resampler: invalid arguments");
  }
  ResamplerDesign d;
  d.bits = 0;
  while ((1u << d.bits) < npfb) {
    d.bits++;
  }
  d.npfb = 1u << d.bits;
  d.subLen = 2 * m;
  d.step = resamplerStep(rate);
  const unsigned n = 2 * m * d.npfb + 1;
  const std::vector<float> proto = kaiserLowpass(n, fc / static_cast<float>(d.npfb), As, 0.0f);
  float total = 0.0f;
  for (unsigned i = 0; i < n; i++) {
    total += proto[i];
  }
  const float gain = static_cast<float>(d.npfb) / total;
  d.bank.assign(static_cast<size_t>(d.npfb) * d.subLen, 0.0f);
  for (unsigned b = 0; b < d.npfb; b++) {
    for (unsigned k = 0; k < d.subLen; k++) {
      // tap k of branch b multiplies x[now-k]; window order puts it at subLen-1-k
      d.bank[static_cast<size_t>(b) * d.subLen + (d.subLen - 1 - k)] = proto[b + k * d.npfb] * gain;
    }
  }
  return d;
}

SymSyncDesign symsyncRrc(unsigned k, unsigned m, float beta, unsigned npfb, float bt) {
  SymSyncDesign d;
  d.k = k;
  d.npfb = npfb;
  const std::vector<float> H = rootRaisedCosine(k * npfb, m, beta);
  const unsigned len = static_cast<unsigned>(H.size());
  std::vector<float> dH(len);
  float peak = 0.0f;
  for (unsigned i = 0; i < len; i++) {
    const float next = (i == len - 1) ? H[0] : H[i + 1];
    const float prev = (i == 0) ? H[len - 1] : H[i - 1];
    dH[i] = next - prev;
    const float p = std::fabs(H[i] * dH[i]);
    if (p > peak || i == 0) {
      peak = p;
    }
  }
  for (unsigned i = 0; i < len; i++) {
    dH[i] = dH[i] * (0.06f / peak);
  }
  d.subLen = len / npfb;
  d.mf.assign(static_cast<size_t>(npfb) * d.subLen, 0.0f);
  d.dmf.assign(static_cast<size_t>(npfb) * d.subLen, 0.0f);
  for (unsigned b = 0; b < npfb; b++) {
    for (unsigned j = 0; j < d.subLen; j++) {
      d.mf[static_cast<size_t>(b) * d.subLen + (d.subLen - 1 - j)] = H[b + j * npfb];
      d.dmf[static_cast<size_t>(b) * d.subLen + (d.subLen - 1 - j)] = dH[b + j * npfb];
    }
  }
  const float alpha = 1.000f - bt;
  const float betaLf = 0.220f * bt;
  const float A0 = 1.00f - 0.500f * alpha;
  const float A1 = -0.495f * alpha;
  d.sosB0 = betaLf / A0;
  d.sosA1 = A1 / A0;
  d.rateAdjustment = 0.5f * bt;
  return d;
}

ChannelFilterSpec channelFilterSpec(int bwHz, int w0Hz, int inputRate) {
  static const int kBw[30] = {309000, 298000, 281000, 263000, 246000, 229000, 211000, 194000,
                              177000, 159000, 142000, 125000, 108000, 95000,  90000,  83000,
                              73000,  63000,  55000,  48000,  42000,  36000,  32000,  27000,
                              24000,  20000,  17000,  15000,  9000,   0};
  ChannelFilterSpec s;
  const int effective = (bwHz <= 0) ? w0Hz : bwHz;
  s.index = 29;
  if (effective > 0) {
    int best = 0x7fffffff;
    for (int i = 0; i < 29; i++) {
      const int diff = std::abs(kBw[i] - effective);
      if (diff < best) {
        best = diff;
        s.index = i;
      }
    }
  }
  const int sel = kBw[s.index];
  const double headroom = 0.45 * static_cast<double>(inputRate);
  const double cutHz = (sel > 0) ? std::clamp(static_cast<double>(sel) * 0.5, 9000.0, headroom)
                                 : headroom;
  s.cutoff = std::clamp(static_cast<float>(cutHz / static_cast<double>(inputRate)), 0.01f, 0.45f);
  s.length = (sel > 0 && sel <= 73000) ? 121u : 81u;
  s.atten = (sel > 0 && sel <= 42000) ? 70.0f : 60.0f;
  return s;
}

int tefBandwidthHz(int mode) {
  static const int kTef[17] = {311000, 287000, 254000, 236000, 217000, 200000, 184000, 168000, 151000,
                               133000, 114000, 97000,  84000,  72000,  64000,  56000,  0};
  return kTef[std::clamp(mode, 0, 16)];
}

uint32_t ncoConstrain(float radians) {
  const float p = static_cast<float>(static_cast<double>(radians) * 0.159154943091895);
  float frac = p - static_cast<float>(static_cast<long>(p));
  if (frac < 0.0f) {
    frac = frac + 1.0f;
  }
  const float scaled = frac * 4294967296.0f;
  return static_cast<uint32_t>(static_cast<uint64_t>(scaled));
}

}  // namespace fmdesign
