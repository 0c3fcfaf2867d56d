/* oracle/liquid_shim/liquid/liquid.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A stand-in for liquid-dsp's public header that declares exactly the C entry points the
 * reference's hot-path sources call (call sites: SURVEY.md §8(c); include site
 * /root/reference/include/dsp/liquid_primitives.h:10-16 and
 * include/redsea_port/dsp/liquid_wrappers.hh:33-40). liquid-dsp itself is absent from this image
 * (un-vendored, un-pinned dependency of the reference). With this header on the include path the
 * reference's OWN sources — src/dsp/liquid_primitives.cpp, src/fm_demod.cpp,
 * src/stereo_decoder.cpp, src/af_post_processor.cpp, src/rds_decoder.cpp, src/redsea_port/** —
 * compile UNMODIFIED, in place, into oracle/_ref/libfmref.so (oracle/Makefile); the definitions
 * are in liquid_shim.cpp, over the restatement of liquid's published algorithms in
 * ../liquid_restated.hpp. What stays unpinned after that is liquid's internals only.
 *
 * Signatures follow liquid-dsp >= 1.4 (objects are opaque pointers, functions return int).
 */
#ifndef ORACLE_LIQUID_SHIM_LIQUID_H_
#define ORACLE_LIQUID_SHIM_LIQUID_H_

#ifdef __cplusplus
#include <complex>
typedef std::complex<float> liquid_float_complex;
extern "C" {
#else
#include <complex.h>
typedef float complex liquid_float_complex;
#endif

/* the reference switches on this for liquid_firdes_kaiser's return type
 * (src/dsp/liquid_primitives.cpp:8-15); the shim returns int and is called as a statement */
#define LIQUID_VERSION "1.6.0"
#define LIQUID_VERSION_NUMBER 1006000

#define LIQUID_OK 0

typedef enum { LIQUID_NCO = 0, LIQUID_VCO } liquid_ncotype;

typedef enum {
  LIQUID_FIRFILT_UNKNOWN = 0,
  LIQUID_FIRFILT_KAISER,
  LIQUID_FIRFILT_PM,
  LIQUID_FIRFILT_RCOS,
  LIQUID_FIRFILT_FEXP,
  LIQUID_FIRFILT_FSECH,
  LIQUID_FIRFILT_FARCSECH,
  LIQUID_FIRFILT_ARKAISER,
  LIQUID_FIRFILT_RKAISER,
  LIQUID_FIRFILT_RRC
} liquid_firfilt_type;

typedef enum { LIQUID_MODEM_UNKNOWN = 0, LIQUID_MODEM_PSK2 } modulation_scheme;

typedef struct agc_crcf_s *agc_crcf;
typedef struct firfilt_crcf_s *firfilt_crcf;
typedef struct nco_crcf_s *nco_crcf;
typedef struct freqdem_s *freqdem;
typedef struct iirfilt_rrrf_s *iirfilt_rrrf;
typedef struct resamp_rrrf_s *resamp_rrrf;
typedef struct firdecim_crcf_s *firdecim_crcf;
typedef struct symsync_crcf_s *symsync_crcf;
typedef struct modemcf_s *modemcf;
typedef modemcf modem;

/* filter design */
int liquid_firdes_kaiser(unsigned int n, float fc, float As, float mu, float *h);

/* agc_crcf */
agc_crcf agc_crcf_create(void);
int agc_crcf_destroy(agc_crcf q);
int agc_crcf_set_bandwidth(agc_crcf q, float bt);
int agc_crcf_set_gain(agc_crcf q, float gain);
int agc_crcf_execute(agc_crcf q, liquid_float_complex x, liquid_float_complex *y);

/* firfilt_crcf */
firfilt_crcf firfilt_crcf_create(float *h, unsigned int n);
firfilt_crcf firfilt_crcf_create_kaiser(unsigned int n, float fc, float As, float mu);
int firfilt_crcf_destroy(firfilt_crcf q);
int firfilt_crcf_set_scale(firfilt_crcf q, float scale);
int firfilt_crcf_push(firfilt_crcf q, liquid_float_complex x);
int firfilt_crcf_execute(firfilt_crcf q, liquid_float_complex *y);
unsigned int firfilt_crcf_get_length(firfilt_crcf q);
float firfilt_crcf_groupdelay(firfilt_crcf q, float fc);

/* nco_crcf */
nco_crcf nco_crcf_create(liquid_ncotype type);
int nco_crcf_destroy(nco_crcf q);
int nco_crcf_reset(nco_crcf q);
int nco_crcf_set_frequency(nco_crcf q, float dtheta);
int nco_crcf_step(nco_crcf q);
float nco_crcf_get_phase(nco_crcf q);
int nco_crcf_pll_set_bandwidth(nco_crcf q, float bw);
int nco_crcf_pll_step(nco_crcf q, float dphi);

/* freqdem */
freqdem freqdem_create(float kf);
int freqdem_destroy(freqdem q);
int freqdem_reset(freqdem q);
int freqdem_demodulate(freqdem q, liquid_float_complex r, float *m);

/* iirfilt_rrrf */
iirfilt_rrrf iirfilt_rrrf_create(float *b, unsigned int nb, float *a, unsigned int na);
iirfilt_rrrf iirfilt_rrrf_create_dc_blocker(float alpha);
int iirfilt_rrrf_destroy(iirfilt_rrrf q);
int iirfilt_rrrf_execute(iirfilt_rrrf q, float x, float *y);

/* resamp_rrrf */
resamp_rrrf resamp_rrrf_create(float rate, unsigned int m, float fc, float As, unsigned int npfb);
int resamp_rrrf_destroy(resamp_rrrf q);
int resamp_rrrf_set_rate(resamp_rrrf q, float rate);
int resamp_rrrf_execute(resamp_rrrf q, float x, float *y, unsigned int *num_written);

/* firdecim_crcf */
firdecim_crcf firdecim_crcf_create(unsigned int M, float *h, unsigned int h_len);
int firdecim_crcf_destroy(firdecim_crcf q);
int firdecim_crcf_set_scale(firdecim_crcf q, float scale);
int firdecim_crcf_execute(firdecim_crcf q, liquid_float_complex *x, liquid_float_complex *y);

/* symsync_crcf */
symsync_crcf symsync_crcf_create_rnyquist(int type, unsigned int k, unsigned int m, float beta,
                                          unsigned int M);
int symsync_crcf_destroy(symsync_crcf q);
int symsync_crcf_reset(symsync_crcf q);
int symsync_crcf_set_lf_bw(symsync_crcf q, float bt);
int symsync_crcf_set_output_rate(symsync_crcf q, unsigned int k_out);
int symsync_crcf_execute(symsync_crcf q, liquid_float_complex *x, unsigned int nx,
                         liquid_float_complex *y, unsigned int *ny);

/* modem (BPSK only) — both spellings (include/redsea_port/dsp/liquid_wrappers.hh:146-150) */
modemcf modemcf_create(modulation_scheme scheme);
int modemcf_destroy(modemcf q);
int modemcf_demodulate(modemcf q, liquid_float_complex x, unsigned int *s);
float modemcf_get_demodulator_phase_error(modemcf q);
modem modem_create(modulation_scheme scheme);
int modem_destroy(modem q);
int modem_demodulate(modem q, liquid_float_complex x, unsigned int *s);
float modem_get_demodulator_phase_error(modem q);

#ifdef __cplusplus
}
#endif
#endif /* ORACLE_LIQUID_SHIM_LIQUID_H_ */
