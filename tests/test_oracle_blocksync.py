"""The integer RDS back end: oracle restatement vs the REFERENCE's own block_sync.cpp.

oracle/_ref/libredsea_ref.so is built by oracle/Makefile from the reference sources where they
lie (src/redsea_port/{block_sync,group,util/util}.cpp); this is the one part of the hot path
whose parity is pinned against real reference code. Constants are also checked against the
values in IEC 62106 Annex B as quoted in SURVEY.md §8(c).
"""
import numpy as np
import pytest

from oracle import orc
from tests.common import groups_equal


def _stream(rng, n_rep=20):
    groups = orc.rds_groups_ps_rt(0xABCD, "TESTPS  ", "hello world radio text") + \
        [(0x1234, 0x0800 | (15 << 12), 0x1234, 0x5678, 1)]  # a version-B group (C' offset)
    return np.tile(orc.rds_encode_groups(groups), n_rep)


def test_offset_words_and_syndromes(orc_libm):
    syn = orc_libm.lib.orc_rds_syndrome
    # offset word -> syndrome, IEC 62106 Table B.1 (block_sync.cpp:73-77,139-143)
    table = {0x0FC: 0x3D8, 0x198: 0x3D4, 0x168: 0x25C, 0x350: 0x3CC, 0x1B4: 0x258}
    for word, s in table.items():
        assert syn(word) == s
    # any valid codeword has syndrome == syndrome(offset word)
    lib = orc._siglib()
    for data in (0x0000, 0xFFFF, 0x1234, 0xE0CD):
        for idx, word in enumerate((0x0FC, 0x198, 0x168, 0x350, 0x1B4)):
            assert syn(lib.sig_rds_encode_block(data, idx)) == table[word]


def test_clean_stream_roundtrip(orc_libm):
    groups = orc.rds_groups_ps_rt(0x4321, "ROUNDTRP", "round trip text")
    bits = np.tile(orc.rds_encode_groups(groups), 4)
    out = orc_libm.blockstream(bits)
    assert len(out) >= 3 * len(groups)
    pi, ps, rt = orc.decode_ps_rt(out)
    assert (pi, ps, rt) == (0x4321, "ROUNDTRP", "round trip text")
    # after sync every group is clean
    assert (out["errors"][2:] == 0).all()


@pytest.mark.parametrize("ber", [0.0, 0.001, 0.01, 0.03, 0.1, 0.3])
def test_restatement_matches_reference(orc_libm, ber):
    if orc.ref_blockstream(np.zeros(4, np.uint8)) is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    rng = np.random.default_rng(int(ber * 1000) + 7)
    base = _stream(rng)
    total = 0
    for trial in range(4):
        b = np.concatenate([rng.integers(0, 2, rng.integers(0, 60)).astype(np.uint8), base])
        b ^= (rng.random(b.size) < ber).astype(np.uint8)
        if trial % 2:
            lo = rng.integers(1000, 6000)
            b[lo:lo + rng.integers(100, 3000)] = rng.integers(0, 2)  # drop-out: forces sync loss
        mine = orc_libm.blockstream(b)
        ref = orc.ref_blockstream(b)
        assert groups_equal(mine, ref)
        total += len(ref)
    if ber <= 0.03:
        assert total > 100


def test_single_and_double_bit_fec(orc_libm):
    groups = orc.rds_groups_ps_rt(0x7777, "FECTEST ")
    bits = np.tile(orc.rds_encode_groups(groups), 6)
    clean = orc_libm.blockstream(bits)
    b = bits.copy()
    # after sync is established flip one bit in one block and a 2-bit burst in another
    b[104 * 8 + 5] ^= 1
    b[104 * 10 + 40] ^= 1
    b[104 * 10 + 41] ^= 1
    out = orc_libm.blockstream(b)
    assert len(out) == len(clean)
    for k in ("a", "b", "c", "d"):
        assert (out[k] == clean[k]).all()          # data corrected
    assert (out["errors"] != clean["errors"]).sum() == 2  # and flagged as corrected
    if orc.ref_blockstream(b) is not None:
        assert groups_equal(out, orc.ref_blockstream(b))
