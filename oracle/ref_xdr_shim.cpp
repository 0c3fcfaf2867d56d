// oracle/ref_xdr_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Drives the REFERENCE's own XDRServer::updateRDS (src/xdr_server.cpp:403-457), compiled in place by
// oracle/Makefile into oracle/_ref/libxdr_ref.so, and hands back the lines it queued. The queue is
// a private member: the header is included with private access opened for this test shim only.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>

#define private public
#include "xdr_server.h"
#undef private

extern "C" {

void *ref_xdr_create() { return new XDRServer(0); }

// what start / retune do to the PI history through the public interface (xdr_server.cpp:461-470)
void ref_xdr_retune(void *h) { static_cast<XDRServer *>(h)->setFrequencyState(98500000); }

void ref_xdr_destroy(void *h) { delete static_cast<XDRServer *>(h); }

// feeds one group; copies the lines queued by this call (at most `cap`, 32 bytes each)
int ref_xdr_update(void *h, uint16_t a, uint16_t b, uint16_t c, uint16_t d, uint8_t errors,
                   char *lines, int cap) {
  XDRServer *s = static_cast<XDRServer *>(h);
  const size_t before = s->m_rdsQueue.size();
  s->updateRDS(a, b, c, d, errors);
  int n = 0;
  for (size_t i = before; i < s->m_rdsQueue.size() && n < cap; i++, n++) {
    std::strncpy(lines + 32 * n, s->m_rdsQueue[i].second.c_str(), 31);
    lines[32 * n + 31] = 0;
  }
  if (s->m_rdsQueue.size() > 128) {
    s->m_rdsQueue.clear();
  }
  return n;
}

}  // extern "C"
