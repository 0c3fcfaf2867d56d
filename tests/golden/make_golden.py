"""Regenerate tests/golden/*.npz.

Two kinds of fixtures (the reference ships no golden vectors for this path):

* `config1_240k.npz`, `config1_256k.npz`: regression pins of the restated oracle in its `fm` flavour
  (run-time arithmetic independent of the host libm): the flavour the engine is bit-exact against.
* `reference_config1_240k.npz`, `reference_config3_ch7.npz` (round 2): OUTPUTS OF THE REFERENCE'S OWN
  SOURCES run in this container — `oracle/_ref/libfmref.so`, i.e. the unmodified fm_demod / stereo_decoder
  / af_post_processor / rds_decoder / liquid_primitives / redsea_port compiled in place over
  oracle/liquid_shim (oracle/Makefile). They pin the restated oracle (libm flavour) to the reference's
  code wherever /root/reference is absent. liquid-dsp's own internals stay a restatement
  (oracle/liquid_restated.hpp); nothing here comes from a real liquid-dsp build.

Run:  python tests/golden/make_golden.py   (needs /root/reference for the second kind)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "config1_240k": dict(iq_rate=2_400_000, decimation=10),
    "config1_256k": dict(iq_rate=2_048_000, decimation=8),
}
NBLK = 12


def main():
    lib = orc.OracleLib("fm")
    for name, kw in CASES.items():
        sig = orc.config1_signal(fs_iq=kw["iq_rate"])
        iq = sig.generate(NBLK * 8192 * kw["decimation"])
        ch = orc.Channel(lib, orc.make_config(**kw))
        r = ch.process(iq, debug=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            iq_sha256=np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
            iq_head=iq[:64],
            left=r.left, right=r.right,
            mpx_every64=r.mpx[::64],
            status=r.status, groups=r.groups, rds_bits=ch.rds_bits())
        print(name, "audio", r.left.size, "groups", len(r.groups), "bits", ch.rds_bits().size)


REF_CASES = {
    "reference_config1_240k": ("config1", None, 14),
    "reference_config3_ch7": ("config3", 7, 14),
}


def main_reference():
    if not orc.OracleLib.have_ref("ref"):
        print("oracle/_ref/libfmref.so is not built (no /root/reference): reference fixtures unchanged")
        return
    lib = orc.OracleLib("ref")
    kw = dict(iq_rate=2_400_000, decimation=10)
    for name, (kind, c, nblk) in REF_CASES.items():
        sig = orc.config1_signal(fs_iq=kw["iq_rate"]) if kind == "config1" else \
            orc.config3_signal(c, fs_iq=kw["iq_rate"])
        iq = sig.generate(nblk * 8192 * kw["decimation"])
        ch = orc.Channel(lib, orc.make_config(**kw))
        ch.enable_bits_tap()
        r = ch.process(iq, debug=True)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            iq_sha256=np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
            left=r.left, right=r.right, mpx_every16=r.mpx[::16], dec_every64=r.dec[::64],
            status=r.status, groups=r.groups, rds_bits=ch.rds_bits())
        print(name, "audio", r.left.size, "groups", len(r.groups), "bits", ch.rds_bits().size)


if __name__ == "__main__":
    main()
    main_reference()
