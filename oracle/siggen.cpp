// oracle/siggen.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Synthetic FM-stereo + RDS multiplex -> uint8 IQ generator.
// The reference ships no signal generator and no IQ fixtures; this is
// the common input source for the oracle and the engine in tests/ (the engine has
// its own on-device generator for bench.py, fmtuner_sdr_b200/csrc/synth.cu).
//
//   m(t)  = g*(L+R) + g*(L-R)*sin(2 wp t) + a_pilot*sin(wp t) + a_rds*d(t)*sin(3 wp t)
//   phi   = 2 pi * dev * integral(m) + 2 pi * f_off * t
//   iq    = A*exp(j phi) + dc + awgn ;  u8 = clamp(round(127.5 + 127.5*x), 0, 255)
//   d(t)  = sum_k c_k * rrc(t/Tc - k), chips c_k at 2375/s: each differentially
//           encoded bit e -> (+1,-1) if e else (-1,+1)   (biphase, IEC 62106)
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

extern "C" {

struct sig_params {
  double fs_iq;        // IQ sample rate
  double deviation;    // peak deviation for |m| = 1 (75000)
  double tone_l_hz, tone_l_amp;
  double tone_r_hz, tone_r_amp;
  double audio_gain;   // 0.43 (applied to L+R and L-R)
  double pilot_amp;    // 0.10 (0 => mono transmission)
  double rds_amp;      // 0.04 (0 => no RDS)
  double iq_amp;       // 0.5 of full scale
  double snr_db;       // carrier / noise power over the full IQ bandwidth; >= 200 => none
  double freq_offset_hz;
  double dc_i, dc_q;
  uint64_t seed;
  uint64_t start_sample;  // generate samples [start, start+n) of the infinite stream
};

}  // extern "C"

namespace {

double rrc(double z, double beta) {  // root raised cosine, T = 1
  if (std::fabs(z) < 1e-9) {
    return 1.0 - beta + 4.0 * beta / M_PI;
  }
  const double g = 1.0 - 16.0 * beta * beta * z * z;
  if (std::fabs(g) < 1e-9) {
    return beta / std::sqrt(2.0) *
           ((1.0 + 2.0 / M_PI) * std::sin(0.25 * M_PI / beta) +
            (1.0 - 2.0 / M_PI) * std::cos(0.25 * M_PI / beta));
  }
  return (std::sin(M_PI * z * (1.0 - beta)) + 4.0 * beta * z * std::cos(M_PI * z * (1.0 + beta))) /
         (M_PI * z * g);
}

struct Rng {  // splitmix64 + Box-Muller
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  double uniform() { return (static_cast<double>(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
  void gauss2(double *a, double *b) {
    const double r = std::sqrt(-2.0 * std::log(uniform()));
    const double t = 2.0 * M_PI * uniform();
    *a = r * std::cos(t);
    *b = r * std::sin(t);
  }
};

}  // namespace

extern "C" {

// bits: RDS data bits (MSB first per block), repeated cyclically; n_bits may be 0.
// Writes n interleaved I,Q byte pairs. Deterministic in (params, bits).
int sig_generate(const sig_params *p, const uint8_t *bits, size_t n_bits, uint8_t *iq, size_t n) {
  if (!p || !iq || p->fs_iq <= 0) {
    return -1;
  }
  constexpr int kSpan = 5;        // pulse support +-5 chips
  constexpr int kOver = 1024;     // table points per chip
  constexpr double kBeta = 0.8;
  static std::vector<float> table;
  if (table.empty()) {
    std::vector<float> t(static_cast<size_t>(2 * kSpan * kOver + 2));
    for (size_t i = 0; i < t.size(); i++) {
      t[i] = static_cast<float>(rrc(static_cast<double>(i) / kOver - kSpan, kBeta));
    }
    table.swap(t);
  }
  // chips for one cycle of the bit pattern (differential state carried around the
  // cycle: the cycle length is doubled if the parity of ones is odd so it closes)
  std::vector<int8_t> chips;
  if (n_bits > 0 && p->rds_amp != 0.0) {
    int ones = 0;
    for (size_t i = 0; i < n_bits; i++) {
      ones += bits[i] & 1;
    }
    const size_t reps = (ones % 2) ? 2 : 1;
    int e = 0;
    chips.reserve(2 * n_bits * reps);
    for (size_t r = 0; r < reps; r++) {
      for (size_t i = 0; i < n_bits; i++) {
        e ^= (bits[i] & 1);
        chips.push_back(e ? 1 : -1);
        chips.push_back(e ? -1 : 1);
      }
    }
  }
  const long n_chips = static_cast<long>(chips.size());
  const double chip_rate = 2375.0;
  const double wp = 2.0 * M_PI * 19000.0;
  const double dt = 1.0 / p->fs_iq;
  const double sigma = (p->snr_db >= 200.0)
                           ? 0.0
                           : p->iq_amp * std::sqrt(0.5 / std::pow(10.0, p->snr_db / 10.0));
  Rng rng{p->seed * 0x2545f4914f6cdd1dULL + 0x1234567ULL};
  // phase integral restarted per call from an analytic value is not possible for the
  // RDS term, so integrate from sample 0 of the stream when start_sample > 0.
  double phi = 0.0;
  const uint64_t first = p->start_sample;
  const uint64_t last = p->start_sample + n;
  for (uint64_t k = 0; k < last; k++) {
    const double t = static_cast<double>(k) * dt;
    const double l = p->tone_l_amp * std::sin(2.0 * M_PI * p->tone_l_hz * t);
    const double r = p->tone_r_amp * std::sin(2.0 * M_PI * p->tone_r_hz * t);
    double mm = p->audio_gain * (l + r) + p->audio_gain * (l - r) * std::sin(2.0 * wp * t) +
                p->pilot_amp * std::sin(wp * t);
    if (n_chips > 0) {
      const double u = t * chip_rate;
      const long c0 = static_cast<long>(std::floor(u));
      double d = 0.0;
      for (int j = -kSpan + 1; j <= kSpan; j++) {
        const long ci = c0 + j;
        const double z = u - static_cast<double>(ci);  // in (-kSpan, kSpan]
        const double pos = (z + kSpan) * kOver;
        const long ip = static_cast<long>(pos);
        const double fr = pos - static_cast<double>(ip);
        const double pv = table[static_cast<size_t>(ip)] * (1.0 - fr) +
                          table[static_cast<size_t>(ip) + 1] * fr;
        long cm = ci % n_chips;
        if (cm < 0) {
          cm += n_chips;
        }
        d += chips[static_cast<size_t>(cm)] * pv;
      }
      mm += p->rds_amp * d * std::sin(3.0 * wp * t);
    }
    phi += 2.0 * M_PI * (p->deviation * mm + p->freq_offset_hz) * dt;
    if (phi > M_PI) {
      phi -= 2.0 * M_PI;
    } else if (phi < -M_PI) {
      phi += 2.0 * M_PI;
    }
    double n1 = 0.0, n2 = 0.0;
    if (sigma > 0.0) {
      rng.gauss2(&n1, &n2);
    }
    if (k < first) {
      continue;
    }
    const double xi = p->iq_amp * std::cos(phi) + p->dc_i + sigma * n1;
    const double xq = p->iq_amp * std::sin(phi) + p->dc_q + sigma * n2;
    const double bi = std::nearbyint(127.5 + 127.5 * xi);
    const double bq = std::nearbyint(127.5 + 127.5 * xq);
    const size_t o = static_cast<size_t>(k - first) * 2;
    iq[o] = static_cast<uint8_t>(bi < 0.0 ? 0.0 : (bi > 255.0 ? 255.0 : bi));
    iq[o + 1] = static_cast<uint8_t>(bq < 0.0 ? 0.0 : (bq > 255.0 ? 255.0 : bq));
  }
  return 0;
}

// RDS block encoder: 16 data bits + (CRC10 over g(x)=x^10+x^8+x^7+x^5+x^4+x^3+1) ^ offset.
// offset index: 0 A, 1 B, 2 C, 3 C', 4 D (values as in src/redsea_port/block_sync.cpp:139-143).
uint32_t sig_rds_encode_block(uint16_t data, int offset_index) {
  static const uint32_t offs[5] = {0x0FC, 0x198, 0x168, 0x350, 0x1B4};
  uint32_t reg = static_cast<uint32_t>(data) << 10;
  for (int i = 25; i >= 10; i--) {
    if (reg & (1u << i)) {
      reg ^= (0x5B9u << (i - 10));
    }
  }
  const uint32_t check = (reg & 0x3FFu) ^ offs[offset_index];
  return (static_cast<uint32_t>(data) << 10) | check;
}

}  // extern "C"
