// oracle/liquid_restated.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the liquid-dsp objects the reference hot path calls
// (SURVEY.md Appendix A). liquid-dsp is an external, un-vendored, un-pinned
// dependency of the reference (CMakeLists.txt:122-147) and is absent from this
// image, so its published algorithms are restated here. PARITY UNPINNED against
// a real liquid-dsp build: no golden vectors exist in the reference.
// Everything ABOVE this file is pinned: the reference's own wrapper and class code is
// compiled unmodified over oracle/liquid_shim (which forwards to these objects) and
// compared bit for bit with oracle/pipeline.hpp (tests/test_oracle_vs_reference.py);
// independent float64 models (scipy / numpy, tests/test_oracle_design.py,
// tests/test_oracle_pipeline.py) check the structure of what is restated here.
//
// Reference call sites restated (all under /root/reference):
//   src/dsp/liquid_primitives.cpp:29-54   agc_crcf
//   src/dsp/liquid_primitives.cpp:73-168  firfilt_crcf (+ liquid_firdes_kaiser :85-90)
//   src/dsp/liquid_primitives.cpp:180-187,287-317 nco_crcf
//   src/dsp/liquid_primitives.cpp:200-219 freqdem
//   src/dsp/liquid_primitives.cpp:237-285 iirfilt_rrrf
//   src/dsp/liquid_primitives.cpp:338-362 resamp_rrrf
//   src/dsp/liquid_primitives.cpp:398-499 firdecim_crcf
//   src/redsea_port/dsp/liquid_wrappers.cpp:149-180 symsync_crcf
//   src/redsea_port/dsp/liquid_wrappers.cpp:182-220 modem (BPSK phase error)
//
// Choices where liquid versions differ ("version-dependent"):
//   * Kaiser window r = 2t/(n-1)                       (liquid >= 1.4)
//   * resamp_rrrf fixed-point 2^24 phase, no branch interpolation (>= 1.4)
//   * nco_crcf uint32 phase / frequency                 (>= 1.3.2)
//   * agc energy smoother evaluated in float
//   * symsync reset clears the MF bank only (the dMF window keeps its samples)
// Arithmetic policy: float32 everywhere liquid uses float; every dot product
// is ONE accumulator, oldest sample first, acc = fma(h, x, acc); everything
// else is unfused (+,-,*,/ each rounded once). Compile with -ffp-contract=off.
// Filter DESIGN (Kaiser, RRC) is evaluated in double and rounded to float.
#ifndef ORACLE_LIQUID_RESTATED_HPP_
#define ORACLE_LIQUID_RESTATED_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#ifdef ORACLE_FM_MATH
// Second oracle build: transcendental kernels shared with the engine so that the
// serial nonlinear loops can be compared bit for bit (see fm_math.h header).
#include "../fmtuner_sdr_b200/csrc/fm_math.h"
#endif

namespace orc {

struct cf32 {
  float re = 0.0f;
  float im = 0.0f;
};

// ---------------------------------------------------------------------------
// math dispatch: libm (faithful to the reference call sites) or fm_math
// ---------------------------------------------------------------------------
namespace m {
#ifdef ORACLE_FM_MATH
inline float sin(float x) { return fm_sinf(x); }
inline float cos(float x) { return fm_cosf(x); }
inline float atan2(float y, float x) { return fm_atan2f(y, x); }
inline float exp(float x) { return fm_expf(x); }
inline float log(float x) { return fm_logf(x); }
inline const char *name() { return "fm_math"; }
#else
inline float sin(float x) { return ::sinf(x); }
inline float cos(float x) { return ::cosf(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float exp(float x) { return ::expf(x); }
inline float log(float x) { return ::logf(x); }
inline const char *name() { return "libm"; }
#endif
}  // namespace m

// ---------------------------------------------------------------------------
// A.1  filter design (design time only; double, rounded to float)
// ---------------------------------------------------------------------------
inline double besseli0(double z) {
  // I0(z) = sum_k ((z/2)^k / k!)^2
  const double h = 0.5 * z;
  double term = 1.0;
  double sum = 1.0;
  for (int k = 1; k < 200; k++) {
    term *= h / static_cast<double>(k);
    const double t2 = term * term;
    sum += t2;
    if (t2 < 1e-22 * sum) {
      break;
    }
  }
  return sum;
}

inline double kaiser_beta_As(double As) {
  As = std::fabs(As);
  if (As > 50.0) {
    return 0.1102 * (As - 8.7);
  }
  if (As > 21.0) {
    return 0.5842 * std::pow(As - 21.0, 0.4) + 0.07886 * (As - 21.0);
  }
  return 0.0;
}

inline double sinc(double x) {
  if (std::fabs(x) < 0.01) {
    return std::cos(M_PI * x / 2.0) * std::cos(M_PI * x / 4.0) * std::cos(M_PI * x / 8.0);
  }
  return std::sin(M_PI * x) / (M_PI * x);
}

inline double kaiser_window(unsigned i, unsigned n, double beta) {
  const double t = static_cast<double>(i) - static_cast<double>(n - 1) / 2.0;
  const double r = 2.0 * t / static_cast<double>(n - 1);
  const double a = besseli0(beta * std::sqrt(std::max(0.0, 1.0 - r * r)));
  const double b = besseli0(beta);
  return a / b;
}

// liquid_firdes_kaiser(n, fc, As, mu, h)
inline std::vector<float> firdes_kaiser(unsigned n, float fc, float As, float mu) {
  if (n == 0 || !(fc > 0.0f) || !(fc <= 0.5f)) {
    throw std::runtime_error("firdes_kaiser: invalid arguments");
  }
  const double beta = kaiser_beta_As(As);
  std::vector<float> h(n);
  for (unsigned i = 0; i < n; i++) {
    const double t = static_cast<double>(i) - static_cast<double>(n - 1) / 2.0 + mu;
    const double h1 = sinc(2.0 * static_cast<double>(fc) * t);
    const double h2 = kaiser_window(i, n, beta);
    h[i] = static_cast<float>(h1 * h2);
  }
  return h;
}

// liquid_firdes_rrcos(k, m, beta, dt, h): 2*k*m+1 taps
inline std::vector<float> firdes_rrcos(unsigned k, unsigned mm, float beta_f, float dt) {
  const unsigned n = 2 * k * mm + 1;
  const double beta = beta_f;
  std::vector<float> h(n);
  for (unsigned i = 0; i < n; i++) {
    const double z = (static_cast<double>(i) + dt) / static_cast<double>(k) - static_cast<double>(mm);
    const double t1 = std::cos((1.0 + beta) * M_PI * z);
    const double t2 = std::sin((1.0 - beta) * M_PI * z);
    double v;
    if (std::fabs(z) < 1e-5) {
      v = 1.0 - beta + 4.0 * beta / M_PI;
    } else {
      const double t3 = 1.0 / (4.0 * beta * z);
      double g = 1.0 - 16.0 * beta * beta * z * z;
      const double t4 = 4.0 * beta / (M_PI * g);
      g *= g;
      if (g < 1e-5) {
        const double g1 = 1.0 + 2.0 / M_PI;
        const double g2 = std::sin(0.25 * M_PI / beta);
        const double g3 = 1.0 - 2.0 / M_PI;
        const double g4 = std::cos(0.25 * M_PI / beta);
        v = beta / std::sqrt(2.0) * (g1 * g2 + g3 * g4);
      } else {
        v = t4 * (t1 + (t2 * t3));
      }
    }
    h[i] = static_cast<float>(v);
  }
  return h;
}

// ---------------------------------------------------------------------------
// dot products: one accumulator, window order (oldest first), fused mul-add
// ---------------------------------------------------------------------------
inline float dot_r(const float *hrev, const float *w, unsigned n) {
  float acc = 0.0f;
  for (unsigned i = 0; i < n; i++) {
    acc = std::fmaf(hrev[i], w[i], acc);
  }
  return acc;
}

inline cf32 dot_c(const float *hrev, const cf32 *w, unsigned n) {
  float ar = 0.0f;
  float ai = 0.0f;
  for (unsigned i = 0; i < n; i++) {
    ar = std::fmaf(hrev[i], w[i].re, ar);
    ai = std::fmaf(hrev[i], w[i].im, ai);
  }
  return cf32{ar, ai};
}

// sliding window with a doubled buffer: view() is contiguous, oldest first
template <typename T> class Window {
public:
  void init(unsigned n) {
    n_ = n;
    buf_.assign(static_cast<size_t>(2) * n, T{});
    pos_ = 0;
  }
  void reset() {
    std::fill(buf_.begin(), buf_.end(), T{});
    pos_ = 0;
  }
  void push(T x) {
    buf_[pos_] = x;
    buf_[pos_ + n_] = x;
    pos_++;
    if (pos_ == n_) {
      pos_ = 0;
    }
  }
  const T *view() const { return buf_.data() + pos_; }
  unsigned size() const { return n_; }

private:
  unsigned n_ = 0;
  unsigned pos_ = 0;
  std::vector<T> buf_;
};

// ---------------------------------------------------------------------------
// A.2  firfilt (complex data / real taps, and a real-data view of the same)
// ---------------------------------------------------------------------------
class FirFiltC {
public:
  void create(const std::vector<float> &h, float scale) {
    h_ = h;
    hrev_.assign(h.rbegin(), h.rend());
    scale_ = scale;
    w_.init(static_cast<unsigned>(h.size()));
  }
  // firfilt_crcf_create_kaiser + set_scale(2 fc)   (liquid_primitives.cpp:73-80)
  void create_kaiser(unsigned n, float fc, float As, float mu) {
    create(firdes_kaiser(n, fc, As, mu), 2.0f * fc);
  }
  void reset() { w_.reset(); }
  void set_scale(float scale) { scale_ = scale; }  // firfilt_crcf_set_scale
  void push(cf32 x) { w_.push(x); }
  cf32 execute() const {
    cf32 y = dot_c(hrev_.data(), w_.view(), w_.size());
    y.re = y.re * scale_;
    y.im = y.im * scale_;
    return y;
  }
  const std::vector<float> &taps() const { return h_; }
  float scale() const { return scale_; }
  unsigned length() const { return w_.size(); }

private:
  std::vector<float> h_, hrev_;
  float scale_ = 1.0f;
  Window<cf32> w_;
};

// Real-input use of firfilt_crcf: the reference pushes complex(x, 0) and keeps
// .real() (stereo_decoder.cpp:172-173,233-236); the real part of liquid's
// complex dot product never sees the imaginary lane, so only it is evaluated.
class FirFiltR {
public:
  void create(const std::vector<float> &h, float scale) {
    h_ = h;
    hrev_.assign(h.rbegin(), h.rend());
    scale_ = scale;
    w_.init(static_cast<unsigned>(h.size()));
  }
  void create_kaiser(unsigned n, float fc, float As, float mu) {
    create(firdes_kaiser(n, fc, As, mu), 2.0f * fc);
  }
  void reset() { w_.reset(); }
  void push(float x) { w_.push(x); }
  float execute() const { return dot_r(hrev_.data(), w_.view(), w_.size()) * scale_; }
  const std::vector<float> &taps() const { return h_; }
  float scale() const { return scale_; }
  unsigned length() const { return w_.size(); }

private:
  std::vector<float> h_, hrev_;
  float scale_ = 1.0f;
  Window<float> w_;
};

// ---------------------------------------------------------------------------
// A.7  firdecim_crcf: output computed right after pushing sample 0 of each M
// ---------------------------------------------------------------------------
class FirDecimC {
public:
  void create(unsigned M, const std::vector<float> &h, float scale) {
    M_ = M;
    h_ = h;
    hrev_.assign(h.rbegin(), h.rend());
    scale_ = scale;
    w_.init(static_cast<unsigned>(h.size()));
  }
  void reset() { w_.reset(); }
  void set_scale(float scale) { scale_ = scale; }  // firdecim_crcf_set_scale
  cf32 execute(const cf32 *x) {
    cf32 y{};
    for (unsigned i = 0; i < M_; i++) {
      w_.push(x[i]);
      if (i == 0) {
        y = dot_c(hrev_.data(), w_.view(), w_.size());
        y.re = y.re * scale_;
        y.im = y.im * scale_;
      }
    }
    return y;
  }
  const std::vector<float> &taps() const { return h_; }
  float scale() const { return scale_; }

private:
  unsigned M_ = 1;
  std::vector<float> h_, hrev_;
  float scale_ = 1.0f;
  Window<cf32> w_;
};

// ---------------------------------------------------------------------------
// A.4  iirfilt_rrrf, order 1, direct form II
// ---------------------------------------------------------------------------
class Iir1 {
public:
  // iirfilt_rrrf_create_dc_blocker(alpha): b = {1,-1}, a = {1, -1+alpha}
  void create_dc_blocker(float alpha) {
    dc_ = true;
    b0_ = 1.0f;
    a1_ = -1.0f + alpha;
    v1_ = 0.0f;
  }
  // iirfilt_rrrf_create(b, 1, a, 2) as used for de-emphasis
  void create_b1_a2(float b0, float a0, float a1) {
    dc_ = false;
    b0_ = b0 / a0;
    a1_ = a1 / a0;
    v1_ = 0.0f;
  }
  void reset() { v1_ = 0.0f; }
  float execute(float x) {
    const float v0 = x - (a1_ * v1_);
    const float y = dc_ ? (v0 - v1_) : (b0_ * v0);
    v1_ = v0;
    return y;
  }
  float a1() const { return a1_; }
  float b0() const { return b0_; }

private:
  bool dc_ = true;
  float b0_ = 1.0f;
  float a1_ = 0.0f;
  float v1_ = 0.0f;
};

// ---------------------------------------------------------------------------
// A.5  agc_crcf
// ---------------------------------------------------------------------------
class Agc {
public:
  void create(float bandwidth, float gain) {
    alpha_ = bandwidth;
    g0_ = gain;
    g_ = gain;
    y2_ = 1.0f;
  }
  void reset() {
    g_ = g0_;
    y2_ = 1.0f;
  }
  void set_bandwidth(float bandwidth) { alpha_ = bandwidth; }  // agc_crcf_set_bandwidth
  void set_gain(float gain) { g_ = gain; }                      // agc_crcf_set_gain
  cf32 execute(cf32 x) {
    cf32 y{x.re * g_, x.im * g_};
    const float e = (y.re * y.re) + (y.im * y.im);
    y2_ = ((1.0f - alpha_) * y2_) + (alpha_ * e);
    if (y2_ > 1e-6f) {
      g_ = g_ * m::exp((-0.5f * alpha_) * m::log(y2_));
    }
    if (g_ > 1e6f) {
      g_ = 1e6f;
    }
    return y;  // output scale = 1
  }

private:
  float alpha_ = 0.01f;
  float g0_ = 1.0f;
  float g_ = 1.0f;
  float y2_ = 1.0f;
};

// ---------------------------------------------------------------------------
// A.8  nco_crcf (uint32 phase), PLL alpha = bw, beta = sqrt(bw)
// ---------------------------------------------------------------------------
inline uint32_t nco_constrain(float theta) {
  const float p = static_cast<float>(static_cast<double>(theta) * 0.159154943091895);
  float fpart = p - static_cast<float>(static_cast<long>(p));
  if (fpart < 0.0f) {
    fpart = fpart + 1.0f;
  }
  // (uint32_t)(fpart * 0xffffffff): the constant converts to 2^32 in float; a
  // product that rounds to 2^32 wraps to 0 (x86 cvttss2si + truncation).
  const float scaled = fpart * 4294967296.0f;
  return static_cast<uint32_t>(static_cast<uint64_t>(scaled));
}

class Nco {
public:
  void create(float freq) {
    f0_ = freq;
    theta_ = 0;
    dtheta_ = nco_constrain(freq);
  }
  // nco_crcf_reset + set_frequency(initial)  (liquid_primitives.cpp:287-292)
  void reset() {
    theta_ = 0;
    dtheta_ = nco_constrain(f0_);
  }
  void reset_zero() {  // nco_crcf_reset alone: phase and frequency to zero
    theta_ = 0;
    dtheta_ = 0;
  }
  void set_frequency(float freq) { dtheta_ = nco_constrain(freq); }  // nco_crcf_set_frequency
  void pll_set_bandwidth(float bw) {
    alpha_ = bw;
    beta_ = std::sqrt(bw);
  }
  void step() { theta_ += dtheta_; }
  void pll_step(float dphi) {
    dtheta_ += nco_constrain(dphi * alpha_);
    theta_ += nco_constrain(dphi * beta_);
  }
  float phase() const {
    return static_cast<float>(6.283185307179586 * static_cast<double>(static_cast<float>(theta_)) /
                              4294967296.0);
  }
  uint32_t theta() const { return theta_; }
  uint32_t dtheta() const { return dtheta_; }

private:
  float f0_ = 0.0f;
  uint32_t theta_ = 0;
  uint32_t dtheta_ = 0;
  float alpha_ = 0.1f;
  float beta_ = 0.31622776f;
};

// ---------------------------------------------------------------------------
// A.3  freqdem
// ---------------------------------------------------------------------------
class FreqDem {
public:
  void create(float kf) {
    ref_ = static_cast<float>(1.0 / (2.0 * M_PI * static_cast<double>(kf)));
    prev_ = cf32{};
  }
  void reset() { prev_ = cf32{}; }
  float demodulate(cf32 r) {
    // arg(conj(prev) * r)
    const float re = (prev_.re * r.re) + (prev_.im * r.im);
    const float im = (prev_.re * r.im) - (prev_.im * r.re);
    prev_ = r;
    return m::atan2(im, re) * ref_;
  }
  float ref() const { return ref_; }

private:
  float ref_ = 1.0f;
  cf32 prev_{};
};

// ---------------------------------------------------------------------------
// A.6  resamp_rrrf (fixed-point phase, 2^bits branches, no interpolation)
// ---------------------------------------------------------------------------
class Resamp {
public:
  void create(float rate, unsigned mm, float fc, float As, unsigned npfb) {
    if (!(rate > 0.0f) || mm == 0 || !(fc > 0.0f) || !(fc < 0.5f) || npfb == 0) {
      throw std::runtime_error("resamp: invalid arguments");
    }
    bits_ = 0;
    while ((1u << bits_) < npfb) {
      bits_++;
    }
    npfb_ = 1u << bits_;
    m_ = mm;
    sub_len_ = 2 * mm;
    set_rate(rate);
    const unsigned n = 2 * mm * npfb_ + 1;
    const std::vector<float> hf = firdes_kaiser(n, fc / static_cast<float>(npfb_), As, 0.0f);
    float gain = 0.0f;
    for (unsigned i = 0; i < n; i++) {
      gain += hf[i];
    }
    gain = static_cast<float>(npfb_) / gain;
    proto_.resize(n);
    for (unsigned i = 0; i < n; i++) {
      proto_[i] = hf[i] * gain;
    }
    // firpfb_create(npfb, h, n-1): branch i, tap k = h[i + k*npfb], stored reversed
    bank_.assign(static_cast<size_t>(npfb_) * sub_len_, 0.0f);
    for (unsigned i = 0; i < npfb_; i++) {
      for (unsigned k = 0; k < sub_len_; k++) {
        bank_[static_cast<size_t>(i) * sub_len_ + (sub_len_ - k - 1)] = proto_[i + k * npfb_];
      }
    }
    w_.init(sub_len_);
    phase_ = 0;
  }
  void set_rate(float rate) {
    rate_ = rate;
    // (uint32_t)round((1<<24)/rate): int/float is a float division
    const float q = static_cast<float>(1 << 24) / rate;
    step_ = static_cast<uint32_t>(std::round(static_cast<double>(q)));
  }
  void reset() {
    w_.reset();
    phase_ = 0;
  }
  unsigned execute(float x, float *y) {
    w_.push(x);
    unsigned n = 0;
    while (phase_ < (1u << 24)) {
      const unsigned idx = phase_ >> (24 - bits_);
      y[n++] = dot_r(&bank_[static_cast<size_t>(idx) * sub_len_], w_.view(), sub_len_);
      phase_ += step_;
    }
    phase_ -= (1u << 24);
    return n;
  }
  uint32_t step() const { return step_; }
  uint32_t phase() const { return phase_; }
  unsigned sub_len() const { return sub_len_; }
  unsigned npfb() const { return npfb_; }
  const std::vector<float> &bank() const { return bank_; }  // [npfb][sub_len], window order

private:
  float rate_ = 1.0f;
  unsigned m_ = 0;
  unsigned bits_ = 5;
  unsigned npfb_ = 32;
  unsigned sub_len_ = 0;
  uint32_t step_ = 1u << 24;
  uint32_t phase_ = 0;
  std::vector<float> proto_;
  std::vector<float> bank_;
  Window<float> w_;
};

// ---------------------------------------------------------------------------
// A.9  symsync_crcf, create_rnyquist(RRC, k, m, beta, M), output rate 1
// ---------------------------------------------------------------------------
class SymSync {
public:
  void create_rnyquist_rrc(unsigned k, unsigned mm, float beta, unsigned M) {
    k_ = k;
    npfb_ = M;
    k_out_ = 1;
    const std::vector<float> H = firdes_rrcos(k * M, mm, beta, 0.0f);  // 2*M*k*m+1 taps
    const unsigned H_len = static_cast<unsigned>(H.size());
    std::vector<float> dH(H_len);
    float hdh_max = 0.0f;
    for (unsigned i = 0; i < H_len; i++) {
      if (i == 0) {
        dH[i] = H[i + 1] - H[H_len - 1];
      } else if (i == H_len - 1) {
        dH[i] = H[0] - H[i - 1];
      } else {
        dH[i] = H[i + 1] - H[i - 1];
      }
      if (std::fabs(H[i] * dH[i]) > hdh_max || i == 0) {
        hdh_max = std::fabs(H[i] * dH[i]);
      }
    }
    for (unsigned i = 0; i < H_len; i++) {
      dH[i] = dH[i] * (0.06f / hdh_max);
    }
    sub_len_ = H_len / M;  // firpfb_create(M, h, H_len): integer division
    mf_.assign(static_cast<size_t>(M) * sub_len_, 0.0f);
    dmf_.assign(static_cast<size_t>(M) * sub_len_, 0.0f);
    for (unsigned i = 0; i < M; i++) {
      for (unsigned n = 0; n < sub_len_; n++) {
        mf_[static_cast<size_t>(i) * sub_len_ + (sub_len_ - n - 1)] = H[i + n * M];
        dmf_[static_cast<size_t>(i) * sub_len_ + (sub_len_ - n - 1)] = dH[i + n * M];
      }
    }
    wmf_.init(sub_len_);
    wdmf_.init(sub_len_);
    set_lf_bw(0.01f);
    reset();
  }
  void set_lf_bw(float bt) {
    const float alpha = 1.000f - bt;
    const float beta = 0.220f * bt;
    const float a = 0.500f;
    const float b = 0.495f;
    const float B0 = beta;
    const float A0 = 1.00f - a * alpha;
    const float A1 = -b * alpha;
    // iirfiltsos_set_coefficients: normalise by A0 (B1=B2=A2=0)
    sos_b0_ = B0 / A0;
    sos_a1_ = A1 / A0;
    rate_adjustment_ = 0.5f * bt;
  }
  void reset() {
    wmf_.reset();  // symsync_crcf_reset clears the MF bank only
    rate_ = static_cast<float>(k_) / static_cast<float>(k_out_);
    del_ = rate_;
    b_ = 0;
    tau_ = 0.0f;
    q_hat_ = 0.0f;
    decim_counter_ = 0;
    sos_v1_ = 0.0f;
  }
  // one input sample; writes up to 8 outputs, returns the count
  unsigned step(cf32 x, cf32 *y) {
    wmf_.push(x);
    wdmf_.push(x);
    unsigned n = 0;
    while (b_ < static_cast<int>(npfb_)) {
      const cf32 mf = dot_c(&mf_[static_cast<size_t>(b_) * sub_len_], wmf_.view(), sub_len_);
      if (n < 8) {
        y[n] = cf32{mf.re / static_cast<float>(k_), mf.im / static_cast<float>(k_)};
      }
      if (decim_counter_ == k_out_) {
        decim_counter_ = 0;
        const cf32 dmf = dot_c(&dmf_[static_cast<size_t>(b_) * sub_len_], wdmf_.view(), sub_len_);
        // advance_internal_loop
        float q = (mf.re * dmf.re) + (mf.im * dmf.im);
        if (q > 1.0f) {
          q = 1.0f;
        } else if (q < -1.0f) {
          q = -1.0f;
        }
        const float v0 = q - (sos_a1_ * sos_v1_);
        q_hat_ = sos_b0_ * v0;
        sos_v1_ = v0;
        rate_ = rate_ + (rate_adjustment_ * q_hat_);
        del_ = rate_ + q_hat_;
      }
      decim_counter_++;
      tau_ = tau_ + del_;
      const float bf = tau_ * static_cast<float>(npfb_);
      b_ = static_cast<int>(std::round(bf));
      n++;
    }
    tau_ = tau_ - 1.0f;
    b_ -= static_cast<int>(npfb_);
    return n;
  }
  unsigned sub_len() const { return sub_len_; }
  const std::vector<float> &mf_bank() const { return mf_; }
  const std::vector<float> &dmf_bank() const { return dmf_; }
  float sos_b0() const { return sos_b0_; }
  float sos_a1() const { return sos_a1_; }
  float rate_adjustment() const { return rate_adjustment_; }

private:
  unsigned k_ = 3, k_out_ = 1, npfb_ = 32, sub_len_ = 0;
  std::vector<float> mf_, dmf_;
  Window<cf32> wmf_, wdmf_;
  float rate_ = 3.0f, del_ = 3.0f, tau_ = 0.0f, q_hat_ = 0.0f;
  int b_ = 0;
  unsigned decim_counter_ = 0;
  float sos_b0_ = 0.0f, sos_a1_ = 0.0f, sos_v1_ = 0.0f, rate_adjustment_ = 0.0f;
};

// A.10 modem BPSK: phase error of the last demodulated sample = Im(x * conj(xhat))
inline float bpsk_phase_error(cf32 x) { return (x.re > 0.0f) ? x.im : -x.im; }

}  // namespace orc

#endif  // ORACLE_LIQUID_RESTATED_HPP_
