// Drop-in replacement for the reference's include/rds_decoder.h (RDSGroup and the public
// surface of RDSDecoder, rds_decoder.h:9-26), implemented over the C ABI in include/fmgpu.h.
#ifndef RDS_DECODER_H
#define RDS_DECODER_H

#include <functional>
#include <stddef.h>
#include <stdint.h>

struct fmgpu_engine;

struct RDSGroup {
  uint16_t blockA;
  uint16_t blockB;
  uint16_t blockC;
  uint16_t blockD;
  uint8_t errors;
};

class RDSDecoder {
public:
  explicit RDSDecoder(int inputRate);
  ~RDSDecoder();
  RDSDecoder(const RDSDecoder &) = delete;
  RDSDecoder &operator=(const RDSDecoder &) = delete;

  void reset();
  // onGroup is invoked synchronously, once per decoded group, in order
  void process(const float *mpx, size_t numSamples,
               const std::function<void(const RDSGroup &)> &onGroup);

private:
  fmgpu_engine *engine_;
  int rate_;
};

#endif
