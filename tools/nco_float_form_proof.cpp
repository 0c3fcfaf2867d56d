// tools/nco_float_form_proof.cpp — exhaustive proof that the float-only forms of liquid's
// nco_crcf constrain() / get_phase() used by the device kernels (kernels.cu: ncoConstrainDev,
// ncoPhaseDev) are bit-identical to the float -> double -> float forms the CPU oracle restates
// (oracle/liquid_restated.hpp: nco_constrain, Nco::phase). ~90 s on 8 threads:
//   g++ -O2 -ffp-contract=off -pthread -o /tmp/nco_proof tools/nco_float_form_proof.cpp && /tmp/nco_proof
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

static const double C1 = 0.159154943091895;
static const float C1hi = (float)C1;
static const float C1lo = (float)(C1 - (double)C1hi);
static const double C2 = 6.283185307179586 / 4294967296.0;
static const float C2hi = (float)C2;
static const float C2lo = (float)(C2 - (double)C2hi);

static inline uint32_t finish(float p) {
  float fpart = p - truncf(p);
  if (fpart < 0.0f) fpart = fpart + 1.0f;
  const float scaled = fpart * 4294967296.0f;
  return (scaled >= 4294967296.0f) ? 0u : (uint32_t)scaled;
}
static inline uint32_t constrain_double(float x) { return finish((float)((double)x * C1)); }
static inline uint32_t constrain_float(float x) {
  const float p = x * C1hi;
  const float e = fmaf(x, C1hi, -p);
  return finish(p + fmaf(x, C1lo, e));
}
static inline float phase_double(float t) { return (float)(6.283185307179586 * (double)t / 4294967296.0); }
static inline float phase_float(float t) {
  const float p = t * C2hi;
  const float e = fmaf(t, C2hi, -p);
  return p + fmaf(t, C2lo, e);
}

int main() {
  std::atomic<unsigned long> bad{0}, n{0};
  auto work = [&](uint64_t lo, uint64_t hi) {
    unsigned long b = 0, k = 0;
    for (uint64_t bb = lo; bb < hi; bb++) {
      const uint32_t bits = (uint32_t)bb;
      float x;
      memcpy(&x, &bits, 4);
      if (!std::isfinite(x) || fabsf(x) >= 9.2e18f) continue;  // (long)p is defined below 2^63
      k++;
      if (constrain_double(x) != constrain_float(x)) b++;
    }
    bad += b;
    n += k;
  };
  std::vector<std::thread> th;
  const int T = 8;
  for (int i = 0; i < T; i++) th.emplace_back(work, (uint64_t)i * (1ull << 32) / T, (uint64_t)(i + 1) * (1ull << 32) / T);
  for (auto &t : th) t.join();
  printf("constrain (resulting uint32): %lu finite floats below 2^63 checked, %lu mismatches\n", n.load(), bad.load());
  unsigned long b2 = 0, n2 = 0;
  for (uint32_t bits = 0; bits <= 0x4f800000u; bits++) {  // every float that is an integer in [0, 2^32]
    float t;
    memcpy(&t, &bits, 4);
    if (t != floorf(t)) continue;
    n2++;
    const float a = phase_double(t), c = phase_float(t);
    if (memcmp(&a, &c, 4) != 0) b2++;
  }
  printf("phase: %lu possible (float)theta checked, %lu mismatches\n", n2, b2);
  printf("constants: C1hi=%a C1lo=%a C2hi=%a C2lo=%a\n", C1hi, C1lo, C2hi, C2lo);
  return (bad.load() || b2) ? 1 : 0;
}
