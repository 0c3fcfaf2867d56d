// engine.h — device-visible data layout of the FM stereo + RDS engine.
//
// Layout in HBM (C = channels, n = DSP-rate samples per call = blocks * N):
//   caller IQ       [C][stride]            uint8 I,Q pairs (read once by k_decim)
//   hist_iq         [C][2*H_IQ]            tail of the previous call's IQ bytes (decimator halo)
//   x1              [C][pitch]   float2    decimated IQ
//   x2              [C][H_X2 + pitch] float2  DC-blocked IQ, H_X2 halo in front (channel FIR)
//   ybuf            [C][1 + pitch] float2  channel-filter output; slot 0 = freqdem r_prev
//   mpx             [C][H_MPX + pitch]     discriminator output, halo for pilot FIR / delay / resamplers
//   pilot           [C][pitch]             19 kHz band-pass output
//   lraw, rraw      [C][H_LR + pitch]      matrix output before the 15 kHz low-pass
//   lf, rf          [C][H_LF + pitch]      low-passed L/R at the DSP rate (resampler halo)
//   audio           [C][2][acap]           32 kHz audio
// Every "halo in front" buffer keeps the last H samples of the previous call in
// [0, H) — k_carry moves them there after each call — so FIR kernels never see a
// block boundary. Serial state (IIR, PLL, AGC, RDS loops) lives in per-channel
// structs, one lane per channel.
#ifndef FMGPU_ENGINE_H_
#define FMGPU_ENGINE_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fmgpu.h"

namespace fmgpu {

constexpr int H_IQ = 512;    // uint8 IQ history, samples (multiple of 8)
constexpr int H_X2 = 128;    // >= padded channel filter length - 1
constexpr int H_MPX = 512;   // >= pilot taps - 1 (<= 510), stereo delay, resampler windows
constexpr int H_LR = 128;    // >= audio filter padded length - 1
constexpr int H_LF = 32;     // >= resampler sub-filter length - 1
constexpr int MAX_TAPS = 512;
constexpr int MAX_CHAN_FILTERS = 32;
constexpr int CHAN_TAPS_PITCH = 128;
constexpr int RDS_LPF_LEN = 255;
constexpr int RDS_NACC = 11;
constexpr int RDS_RING = 256;
constexpr int SS_LEN = 18;   // symsync sub-filter length
constexpr int RDS_RS_LEN = 26;
constexpr int AUD_RS_LEN = 24;
constexpr int RDS_HIST = 28;  // RDS resampler window history kept per channel (>= 25, multiple of 4)
constexpr int Y_OFF = 2;      // ybuf: data starts at float2 index 2 (16-byte aligned), r_prev at 1

struct TapsParam {  // passed by value: lives in the kernel-parameter constant bank
  float h[MAX_TAPS];
};

struct ChanParams {
  int filt;           // index into the channel-filter table
  int agc_mode;       // 0 off
  float agc_alpha;
  int blend_mode;     // 0 soft 1 normal 2 aggressive
  int force_mono;
  int force_stereo;
  int deemph_on;      // AFPostProcessor de-emphasis
  float de_b0, de_a1;
  int mono_deemph_on; // FMDemod mono-path de-emphasis
  float mono_de_b0, mono_de_a1;
  int bandwidth_mode; // FMDemod::m_bandwidthMode (table index)
  int w0_hz;
};

struct DemodState {   // FMDemod carried state (fm_demod.h:55-64)
  float dc_i, dc_q;   // DC-blocker v1
  float agc_g, agc_y2;
  float clip_ratio;
  int clipping;
};

struct StereoState {  // StereoDecoder carried state (stereo_decoder.h:29-54)
  uint32_t theta, dtheta;
  float pbm, mm, pilot_i, pilot_q, blend, pll_freq, pilot_mag;
  int stereo, pilot_count, loss_count, pilot_tenths;
};

struct AudioState {   // AFPostProcessor (af_post_processor.h:27-32) + FMDemod mono chain
  float de_v1[2], dc_v1[2];
  uint32_t rs_phase, rs_phase_next;
  uint32_t n_out;     // outputs of the current logical block
  uint32_t out_base;  // frames written by the earlier blocks of the current call
  float mono_de_v1, mono_dc_v1;
  uint32_t mono_phase, mono_phase_next, mono_n_out;
};

// The 171 kHz resampler's bookkeeping lives apart from RdsState: in the block pipeline the resampler
// of block q + 1 runs beside the demodulator of block q (which holds RdsState in registers and
// writes all of it back). n171 / the 171 kHz rows are double-buffered by block parity.
struct RdsRsState {
  uint32_t rs_phase, rs_phase_next;
  uint32_t n171[2];   // 171 kHz samples the block of that parity produces
};
struct RdsRsRef {     // by value in kernel parameters
  RdsRsState *st;
  int par;
};

struct RdsState {     // redsea SubcarrierSet + BlockStream (subcarrier.hh:38-90, block_sync.hh:67-79)
  uint32_t theta, dtheta;
  float prev_f0_phase, phase0;
  uint32_t since_reset;
  uint32_t ring_pos;
  int realign;
  float2 acc[RDS_NACC];
  float agc_g, agc_y2;
  float2 wmf[SS_LEN], wdmf[SS_LEN];
  float tau, rate, del, q_hat, sos_v1;
  int b;
  uint32_t decim_counter;
  float bi_prev_re, bi_even, bi_odd;
  uint32_t bi_clock, bi_polarity;
  int delta_prev;
  // block synchroniser
  uint32_t bitcount, until, reg, bits_since_lost;
  int expected, in_sync, err_ptr;
  unsigned long long err_mask;
  uint16_t cur_data[4];
  uint32_t cur_recv, cur_err;
  uint32_t pulse_pos[4];
  int pulse_off[4];
  uint32_t n_groups;  // groups emitted in the current call
  uint32_t n_bits;    // bits demodulated in the current call
  uint32_t bits_done; // of those, already consumed by the block synchroniser
};

// Engine-wide constants (by value in kernel parameters).
struct EngineConst {
  int C;             // channels
  int N;             // logical block length (DSP-rate samples)
  int M;             // decimation
  int fs;            // DSP rate
  float fsf;         // (float)fs
  // decimator
  int dec_lp;        // padded tap count (multiple of 4*M)
  float dec_scale;
  // pilot / stereo
  int pil_lp;        // padded pilot tap count
  int delay;         // (pilot taps - 1)/2 + 1
  float nominal_pll; // 2*pi*19000/fs
  float pll_min, pll_max;
  float pll_alpha, pll_beta;
  uint32_t pll_dtheta0;
  float blend_attack[3], blend_release[3], gate[3];
  // audio LPF
  int aud_lp;
  float aud_scale;
  // audio / mono resamplers
  uint32_t aud_step;
  float mono_dc_a1;  // -1 + 0.0008
  float dc_a1_iq;    // -1 + 0.0005
  float dc_a1_af;    // -1 + 0.005
  float fd_ref;      // freqdem 1/(2 pi kf)
  // RDS
  uint32_t rds_step;
  float rds_lpf_scale;
  float rds_agc_alpha;
  uint32_t rds_dtheta0;
  float rds_pll_alpha, rds_pll_beta;
  float ss_b0, ss_a1, ss_rate_adj;
  // k_stereo: CTAs b, b + sm_rot, b + 2 sm_rot ... share an SM and rotate their warp roles (0 = off)
  int sm_rot;
};

}  // namespace fmgpu

#endif  // FMGPU_ENGINE_H_
