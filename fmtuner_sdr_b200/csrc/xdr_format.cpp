// xdr_format.cpp — what an XDR-GTK / FM-DX client sees of a channel's RDS stream: the PI debounce
// and the "P..." / "R..." text lines XDRServer::updateRDS queues for every group
// (src/xdr_server.cpp:189-215 evaluatePiState, :403-457 updateRDS). Host-side integer/byte
// logic for batched use (one state per channel); the single-channel application keeps using the
// reference's own XDRServer, which is unchanged.
#include <cstdio>
#include <cstring>

#include "../../include/fmgpu.h"

namespace {

// xdr_server.cpp:189-215: how trustworthy `value` is among the last `fill` PIs
int piState(const fmgpu_xdr_rds_state *s, uint16_t value) {
  int count = 0, correct = 0;
  for (int i = 0; i < s->pi_fill; i++) {
    if (s->pi_buffer[i] == value) {
      count++;
      if ((s->pi_error[i / 8] & (1u << (i % 8))) == 0) {
        correct++;
      }
    }
  }
  if (correct >= 2) {
    return 0;  // correct
  }
  if (count >= 2 && correct) {
    return 1;  // very likely
  }
  if (count >= 3) {
    return 2;  // likely
  }
  if (count == 2 || correct) {
    return 3;  // unlikely
  }
  return 4;    // invalid
}

}  // namespace

extern "C" {

void fmgpu_xdr_rds_init(fmgpu_xdr_rds_state *s) {
  if (!s) {
    return;
  }
  std::memset(s, 0, sizeof(*s));
  // the constructor and every start / retune leave the history empty with the write position at
  // 63, so the first group lands in slot 0 (xdr_server.cpp:257-266, 461-470)
  s->pi_pos = 63;
  s->pi_last_state = 4;
  s->pi_last_value = 0xFFFF;
}

int fmgpu_xdr_rds_lines(fmgpu_xdr_rds_state *s, const fmgpu_rds_group *g, char lines[2][32]) {
  if (!s || !g || !lines) {
    return 0;
  }
  const unsigned errA = (g->errors >> 6) & 3u, errB = (g->errors >> 4) & 3u;
  s->pi_pos = static_cast<uint8_t>((s->pi_pos + 1) % 64);
  s->pi_buffer[s->pi_pos] = g->a;
  const unsigned bit = 1u << (s->pi_pos % 8);
  if (errA != 0) {
    s->pi_error[s->pi_pos / 8] |= static_cast<uint8_t>(bit);
  } else {
    s->pi_error[s->pi_pos / 8] &= static_cast<uint8_t>(~bit);
  }
  if (s->pi_fill < 64) {
    s->pi_fill++;
  }
  const int state = piState(s, g->a);
  int n = 0;
  if (errA != 3 && state <= 1) {  // block A present and the PI debounced
    std::snprintf(lines[n], 32, "P%04X%.*s", g->a, static_cast<int>(errA), "???");
    n++;
    s->pi_last_value = g->a;
  }
  if (errB == 0) {  // PTY/TP/TA/MS live in block B: groups with a damaged block B are dropped
    std::snprintf(lines[n], 32, "R%04X%04X%04X%02X", g->b, g->c, g->d, g->errors);
    n++;
  }
  s->pi_last_state = static_cast<uint8_t>(state);
  return n;
}

}  // extern "C"
