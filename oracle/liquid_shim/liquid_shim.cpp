// oracle/liquid_shim/liquid_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Definitions of the liquid-dsp C entry points declared in liquid/liquid.h, over the restatement
// of liquid's published algorithms in ../liquid_restated.hpp (libm flavour). Linked with the
// reference's own, unmodified sources into oracle/_ref/libfmref.so, so that what the tests pin the
// oracle against is the reference's code, object for object:
//   src/dsp/liquid_primitives.cpp  (fm_tuner::dsp::liquid::AGC/FIRFilter/NCO/FreqDemod/
//                                   IIRFilterReal/Resampler/ComplexDecimator)
//   src/redsea_port/dsp/liquid_wrappers.cpp (liquid::AGC/FIRFilter/NCO/SymSync/Modem/Resampler)
// Each function notes the liquid-dsp source it follows (SURVEY.md Appendix A holds the algorithm
// statements); unsupported parameter combinations return NULL / an error instead of guessing.
#include "liquid/liquid.h"

#include <cmath>
#include <new>
#include <vector>

#include "../liquid_restated.hpp"

using orc::cf32;

struct agc_crcf_s {
  orc::Agc a;
};
struct firfilt_crcf_s {
  orc::FirFiltC f;
};
struct nco_crcf_s {
  orc::Nco n;
};
struct freqdem_s {
  orc::FreqDem d;
};
struct iirfilt_rrrf_s {
  orc::Iir1 f;
};
struct resamp_rrrf_s {
  orc::Resamp r;
};
struct firdecim_crcf_s {
  orc::FirDecimC d;
  unsigned M = 1;
  std::vector<cf32> tmp;
};
struct symsync_crcf_s {
  orc::SymSync s;
};
struct modemcf_s {
  cf32 r{};  // last demodulated sample
};

namespace {

template <typename T, typename F> T *guarded_new(F &&init) {
  T *q = new (std::nothrow) T();
  if (q == nullptr) {
    return nullptr;
  }
  try {
    init(*q);
  } catch (const std::exception &) {
    delete q;
    return nullptr;
  }
  return q;
}

inline cf32 to_cf(liquid_float_complex x) { return cf32{x.real(), x.imag()}; }

}  // namespace

extern "C" {

// src/filter/src/firdes.c: liquid_firdes_kaiser
int liquid_firdes_kaiser(unsigned int n, float fc, float As, float mu, float *h) {
  try {
    const std::vector<float> v = orc::firdes_kaiser(n, fc, As, mu);
    std::copy(v.begin(), v.end(), h);
  } catch (const std::exception &) {
    return 1;
  }
  return LIQUID_OK;
}

// ---- agc_crcf (src/agc/src/agc.proto.c): create = bandwidth 1e-2, gain 1, y2 = 1 ------------
agc_crcf agc_crcf_create(void) {
  return guarded_new<agc_crcf_s>([](agc_crcf_s &q) { q.a.create(1e-2f, 1.0f); });
}
int agc_crcf_destroy(agc_crcf q) {
  delete q;
  return LIQUID_OK;
}
int agc_crcf_set_bandwidth(agc_crcf q, float bt) {
  q->a.set_bandwidth(bt);
  return LIQUID_OK;
}
int agc_crcf_set_gain(agc_crcf q, float gain) {
  q->a.set_gain(gain);
  return LIQUID_OK;
}
int agc_crcf_execute(agc_crcf q, liquid_float_complex x, liquid_float_complex *y) {
  const cf32 r = q->a.execute(to_cf(x));
  *y = liquid_float_complex(r.re, r.im);
  return LIQUID_OK;
}

// ---- firfilt_crcf (src/filter/src/firfilt.proto.c): scale 1 at creation ----------------------
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n) {
  if (h == nullptr || n == 0) {
    return nullptr;
  }
  return guarded_new<firfilt_crcf_s>(
      [&](firfilt_crcf_s &q) { q.f.create(std::vector<float>(h, h + n), 1.0f); });
}
firfilt_crcf firfilt_crcf_create_kaiser(unsigned int n, float fc, float As, float mu) {
  return guarded_new<firfilt_crcf_s>(
      [&](firfilt_crcf_s &q) { q.f.create(orc::firdes_kaiser(n, fc, As, mu), 1.0f); });
}
int firfilt_crcf_destroy(firfilt_crcf q) {
  delete q;
  return LIQUID_OK;
}
int firfilt_crcf_set_scale(firfilt_crcf q, float scale) {
  q->f.set_scale(scale);
  return LIQUID_OK;
}
int firfilt_crcf_push(firfilt_crcf q, liquid_float_complex x) {
  q->f.push(to_cf(x));
  return LIQUID_OK;
}
int firfilt_crcf_execute(firfilt_crcf q, liquid_float_complex *y) {
  const cf32 r = q->f.execute();
  *y = liquid_float_complex(r.re, r.im);
  return LIQUID_OK;
}
unsigned int firfilt_crcf_get_length(firfilt_crcf q) { return q->f.length(); }
// src/filter/src/group_delay.c: fir_group_delay — Re{ sum h[i] e^{j2 pi fc i} i / sum h[i] e^{..} }
float firfilt_crcf_groupdelay(firfilt_crcf q, float fc) {
  const std::vector<float> &h = q->f.taps();
  std::complex<float> t0 = 0.0f;
  std::complex<float> t1 = 0.0f;
  for (unsigned i = 0; i < h.size(); i++) {
    const std::complex<float> e =
        h[i] * std::exp(std::complex<float>(0.0f, 2.0f * static_cast<float>(M_PI) * fc * static_cast<float>(i)));
    t0 += e * static_cast<float>(i);
    t1 += e;
  }
  return (t0 / t1).real();
}

// ---- nco_crcf (src/nco/src/nco.proto.c): uint32 phase; both types share phase/PLL arithmetic --
nco_crcf nco_crcf_create(liquid_ncotype) {
  return guarded_new<nco_crcf_s>([](nco_crcf_s &q) {
    q.n.create(0.0f);
    q.n.pll_set_bandwidth(0.1f);  // nco_crcf_create: pll bandwidth 0.1
  });
}
int nco_crcf_destroy(nco_crcf q) {
  delete q;
  return LIQUID_OK;
}
int nco_crcf_reset(nco_crcf q) {
  q->n.reset_zero();
  return LIQUID_OK;
}
int nco_crcf_set_frequency(nco_crcf q, float dtheta) {
  q->n.set_frequency(dtheta);
  return LIQUID_OK;
}
int nco_crcf_step(nco_crcf q) {
  q->n.step();
  return LIQUID_OK;
}
float nco_crcf_get_phase(nco_crcf q) { return q->n.phase(); }
int nco_crcf_pll_set_bandwidth(nco_crcf q, float bw) {
  if (bw < 0.0f) {
    return 1;
  }
  q->n.pll_set_bandwidth(bw);
  return LIQUID_OK;
}
int nco_crcf_pll_step(nco_crcf q, float dphi) {
  q->n.pll_step(dphi);
  return LIQUID_OK;
}

// ---- freqdem (src/modem/src/freqdem.proto.c) ---------------------------------------------
freqdem freqdem_create(float kf) {
  if (!(kf > 0.0f)) {
    return nullptr;
  }
  return guarded_new<freqdem_s>([&](freqdem_s &q) { q.d.create(kf); });
}
int freqdem_destroy(freqdem q) {
  delete q;
  return LIQUID_OK;
}
int freqdem_reset(freqdem q) {
  q->d.reset();
  return LIQUID_OK;
}
int freqdem_demodulate(freqdem q, liquid_float_complex r, float *m) {
  *m = q->d.demodulate(to_cf(r));
  return LIQUID_OK;
}

// ---- iirfilt_rrrf (src/filter/src/iirfilt.proto.c), first order, direct form II -------------
iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned int nb, float *a, unsigned int na) {
  if (b == nullptr || a == nullptr || nb != 1 || na != 2) {
    return nullptr;  // the reference creates {alpha} / {1, -(1-alpha)} only (fm_demod.cpp:59-61)
  }
  return guarded_new<iirfilt_rrrf_s>([&](iirfilt_rrrf_s &q) { q.f.create_b1_a2(b[0], a[0], a[1]); });
}
iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha) {
  if (!(alpha > 0.0f)) {
    return nullptr;
  }
  return guarded_new<iirfilt_rrrf_s>([&](iirfilt_rrrf_s &q) { q.f.create_dc_blocker(alpha); });
}
int iirfilt_rrrf_destroy(iirfilt_rrrf q) {
  delete q;
  return LIQUID_OK;
}
int iirfilt_rrrf_execute(iirfilt_rrrf q, float x, float *y) {
  *y = q->f.execute(x);
  return LIQUID_OK;
}

// ---- resamp_rrrf (src/filter/src/resamp.fixed.proto.c) -------------------------------------
resamp_rrrf resamp_rrrf_create(float rate, unsigned int m, float fc, float As, unsigned int npfb) {
  return guarded_new<resamp_rrrf_s>([&](resamp_rrrf_s &q) { q.r.create(rate, m, fc, As, npfb); });
}
int resamp_rrrf_destroy(resamp_rrrf q) {
  delete q;
  return LIQUID_OK;
}
int resamp_rrrf_set_rate(resamp_rrrf q, float rate) {
  if (!(rate > 0.0f)) {
    return 1;
  }
  q->r.set_rate(rate);
  return LIQUID_OK;
}
int resamp_rrrf_execute(resamp_rrrf q, float x, float *y, unsigned int *num_written) {
  *num_written = q->r.execute(x, y);
  return LIQUID_OK;
}

// ---- firdecim_crcf (src/filter/src/firdecim.proto.c) ---------------------------------------
firdecim_crcf firdecim_crcf_create(unsigned int M, float *h, unsigned int h_len) {
  if (M == 0 || h == nullptr || h_len == 0) {
    return nullptr;
  }
  return guarded_new<firdecim_crcf_s>([&](firdecim_crcf_s &q) {
    q.M = M;
    q.tmp.resize(M);
    q.d.create(M, std::vector<float>(h, h + h_len), 1.0f);
  });
}
int firdecim_crcf_destroy(firdecim_crcf q) {
  delete q;
  return LIQUID_OK;
}
int firdecim_crcf_set_scale(firdecim_crcf q, float scale) {
  q->d.set_scale(scale);
  return LIQUID_OK;
}
int firdecim_crcf_execute(firdecim_crcf q, liquid_float_complex *x, liquid_float_complex *y) {
  for (unsigned i = 0; i < q->M; i++) {
    q->tmp[i] = to_cf(x[i]);
  }
  const cf32 r = q->d.execute(q->tmp.data());
  *y = liquid_float_complex(r.re, r.im);
  return LIQUID_OK;
}

// ---- symsync_crcf (src/filter/src/symsync.proto.c) -----------------------------------------
symsync_crcf symsync_crcf_create_rnyquist(int type, unsigned int k, unsigned int m, float beta,
                                          unsigned int M) {
  if (type != LIQUID_FIRFILT_RRC || k < 2 || m == 0 || M == 0) {
    return nullptr;
  }
  return guarded_new<symsync_crcf_s>(
      [&](symsync_crcf_s &q) { q.s.create_rnyquist_rrc(k, m, beta, M); });
}
int symsync_crcf_destroy(symsync_crcf q) {
  delete q;
  return LIQUID_OK;
}
int symsync_crcf_reset(symsync_crcf q) {
  q->s.reset();
  return LIQUID_OK;
}
int symsync_crcf_set_lf_bw(symsync_crcf q, float bt) {
  if (bt < 0.0f || bt > 1.0f) {
    return 1;
  }
  q->s.set_lf_bw(bt);
  return LIQUID_OK;
}
int symsync_crcf_set_output_rate(symsync_crcf, unsigned int k_out) {
  return (k_out == 1) ? LIQUID_OK : 1;  // the reference asks for 1 only (subcarrier.cpp:102)
}
int symsync_crcf_execute(symsync_crcf q, liquid_float_complex *x, unsigned int nx,
                         liquid_float_complex *y, unsigned int *ny) {
  unsigned total = 0;
  for (unsigned i = 0; i < nx; i++) {
    cf32 out[8];
    const unsigned n = q->s.step(to_cf(x[i]), out);
    for (unsigned j = 0; j < n && j < 8; j++) {
      y[total++] = liquid_float_complex(out[j].re, out[j].im);
    }
  }
  *ny = total;
  return LIQUID_OK;
}

// ---- modem, BPSK (src/modem/src/modem_bpsk.proto.c, modem_common.proto.c) -------------------
modemcf modemcf_create(modulation_scheme scheme) {
  if (scheme != LIQUID_MODEM_PSK2) {
    return nullptr;
  }
  return guarded_new<modemcf_s>([](modemcf_s &) {});
}
int modemcf_destroy(modemcf q) {
  delete q;
  return LIQUID_OK;
}
int modemcf_demodulate(modemcf q, liquid_float_complex x, unsigned int *s) {
  *s = (x.real() > 0.0f) ? 0u : 1u;
  q->r = to_cf(x);
  return LIQUID_OK;
}
float modemcf_get_demodulator_phase_error(modemcf q) { return orc::bpsk_phase_error(q->r); }
modem modem_create(modulation_scheme scheme) { return modemcf_create(scheme); }
int modem_destroy(modem q) { return modemcf_destroy(q); }
int modem_demodulate(modem q, liquid_float_complex x, unsigned int *s) {
  return modemcf_demodulate(q, x, s);
}
float modem_get_demodulator_phase_error(modem q) { return modemcf_get_demodulator_phase_error(q); }

}  // extern "C"
