"""Regenerate tests/golden/*.npz — regression pins of the ORACLE (not of the reference).

The reference ships no golden vectors for this path and cannot be built here without
liquid-dsp, so these fixtures pin the in-repo oracle (fm flavour, whose
run-time arithmetic does not depend on the host libm) against accidental change; parity with
a real liquid-dsp build stays UNPINNED. Run:  python tests/golden/make_golden.py
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


if __name__ == "__main__":
    main()
