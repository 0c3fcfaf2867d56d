// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" driver over the REFERENCE's own integer RDS back end
// (src/redsea_port/block_sync.cpp, group.cpp, util/util.cpp), which compiles
// without liquid-dsp. oracle/Makefile compiles those sources where they lie under
// /root/reference into oracle/_ref/libredsea_ref.so; nothing is copied. The
// group packing below follows src/rds_decoder.cpp:29-58.
#include <cstddef>
#include <cstdint>

#include "redsea_port/block_sync.hh"
#include "redsea_port/options.hh"

extern "C" {

struct ref_group {
  uint16_t a, b, c, d;
  uint8_t errors;
  uint8_t pad[3];
  uint32_t bit_index;
};

size_t ref_blockstream_run(const uint8_t *bits, size_t n_bits, ref_group *out, size_t cap) {
  redsea::Options options;
  options.use_fec = true;
  redsea::BlockStream stream;
  stream.init(options);
  size_t k = 0;
  for (size_t i = 0; i < n_bits; i++) {
    stream.pushBit(bits[i] != 0);
    if (!stream.hasGroupReady()) {
      continue;
    }
    const redsea::Group g = stream.popGroup();
    auto val = [&](redsea::eBlockNumber b) -> uint16_t { return g.has(b) ? g.get(b) : 0; };
    auto err = [&](redsea::eBlockNumber b) -> uint8_t {
      if (!g.has(b)) {
        return 3;
      }
      return g.hadErrors(b) ? 1 : 0;
    };
    if (k < cap) {
      out[k].a = val(redsea::BLOCK1);
      out[k].b = val(redsea::BLOCK2);
      out[k].c = val(redsea::BLOCK3);
      out[k].d = val(redsea::BLOCK4);
      out[k].errors = static_cast<uint8_t>((err(redsea::BLOCK1) << 6) | (err(redsea::BLOCK2) << 4) |
                                           (err(redsea::BLOCK3) << 2) | err(redsea::BLOCK4));
      out[k].bit_index = static_cast<uint32_t>(i);
    }
    k++;
  }
  return k;
}

}  // extern "C"
