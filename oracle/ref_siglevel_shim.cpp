// oracle/ref_siglevel_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" driver over the REFERENCE's own computeSignalLevel (src/signal_level.cpp:145-203),
// compiled in place by oracle/Makefile into oracle/_ref/libsiglevel_ref.so.
#include <cstddef>
#include <cstdint>

#include "signal_level.h"

extern "C" void ref_compute_signal_level(const uint8_t *iq, size_t samples, int gain_db,
                                         double comp, double bias, double floor_dbfs,
                                         double ceil_dbfs, double *out5) {
  const SignalLevelResult r = computeSignalLevel(iq, samples, gain_db, comp, bias, floor_dbfs, ceil_dbfs);
  out5[0] = r.level120;
  out5[1] = r.dbfs;
  out5[2] = r.compensatedDbfs;
  out5[3] = r.hardClipRatio;
  out5[4] = r.nearClipRatio;
}
