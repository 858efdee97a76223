"""Philox4x32-10: the oracle's numpy restatement and the library's host entry point against the
Random123 known-answer vectors, and the draw mapping's basic properties.  CPU only."""
import numpy as np
import pytest

from oracle import philox_np as P

KAT = [  # (counter, key, expected) from Random123's kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize('ctr,key,expect', KAT)
def test_oracle_philox_known_answers(ctr, key, expect):
    out = P.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
    assert [int(o[0]) for o in out] == expect


@pytest.mark.parametrize('ctr,key,expect', KAT)
def test_library_philox_known_answers(ctr, key, expect):
    from supervillain_b200 import ops
    assert [int(x) for x in ops.philox4x32_10(ctr, key)] == expect


def test_villain_draw_mapping_ranges_and_symmetry():
    d = P.villain_draws(seed=1234, chain=7, sweep=3, N=64, W=2, interval_n=1)
    assert (d['u'] > 0).all() and (d['u'] < 1).all()
    assert (np.abs(d['dphi']) < np.pi).all()
    for arr in (d['dn_fwd'], d['dn_bwd']):
        assert set(np.unique(arr)) <= {-2, 0, 2}
    # roughly uniform trits
    counts = np.array([(d['dn_fwd'] == k).sum() for k in (-2, 0, 2)])
    assert (np.abs(counts / counts.sum() - 1 / 3) < 0.03).all()
    # different chains / sweeps decorrelate
    e = P.villain_draws(seed=1234, chain=8, sweep=3, N=64, W=2)
    assert not np.array_equal(d['u'], e['u'])


def test_worldline_draw_mapping_ranges():
    for mode, interval, allowed in (('joint', 1, {-1, 1}), ('vortex', 2, {-2, -1, 1, 2}), ('coexact', 1, {-1, 1})):
        d = P.worldline_draws(seed=5, chain=0, sweep=0, N=32, mode=mode, interval=interval)
        assert set(np.unique(d['a'])) == allowed
        assert (d['u'] > 0).all() and (d['u'] < 1).all()
        if mode == 'joint':
            assert set(np.unique(d['b'])) == {-1, 0, 1}
