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


@pytest.mark.gpu
@pytest.mark.parametrize('stream_id', [P.STREAM_VILLAIN_REFINE, P.STREAM_WORLDLINE_REFINE, P.STREAM_VILLAIN_LINK_REFINE, P.STREAM_VILLAIN_SITE_REFINE,
                                       P.STREAM_VILLAIN_EXACT_REFINE, P.STREAM_WORLDLINE_VORTEX_REFINE, P.STREAM_WORLDLINE_COEXACT_REFINE])
def test_lazy_uniform_refinement_matches_oracle(stream_id):
    """The refinement branch of the lazily refined uniform is reached with probability 2^-32 per proposal, i.e. never in a
    sweep test.  svb_debug_decide_lazy forces it: acceptance probabilities placed inside the bracket [f, f + 1] 2^-32 of
    the uniform, on both sides of the refined value, must be decided exactly as the oracle's rule
    u = min(fl(f + (e + 1/2) 2^-32) 2^-32, 1 - 2^-53) decides them."""
    import torch
    from supervillain_b200 import _lib
    rng = np.random.default_rng(stream_id)
    n, seed, chain, sweep = 4000, 0x1234567890ABCDEF, 2**33 + 5, 2**34 + 9
    f = rng.integers(0, 2**32, n, dtype=np.uint64)
    f[:8] = [0, 1, 2**32 - 1, 2**32 - 2, 65535, 65536, 2**31, 12345]
    c0 = rng.integers(0, 2**20, n, dtype=np.uint64)
    word = rng.integers(0, 4, n, dtype=np.uint64)
    blk = P.philox_site(seed, chain, sweep, c0, stream_id)
    e = np.choose(word.astype(np.int64), blk)
    u_ref = np.minimum((f.astype(np.float64) + (e.astype(np.float64) + 0.5) * 2.0**-32) * 2.0**-32, 1.0 - 2.0**-53)
    # A: just above / just below / equal to the refined uniform (inside the bracket), and the bracket's ends and beyond
    kinds = rng.integers(0, 7, n)
    A = np.select([kinds == 0, kinds == 1, kinds == 2, kinds == 3, kinds == 4, kinds == 5],
                  [np.nextafter(u_ref, 2.0), np.nextafter(u_ref, -1.0), u_ref, f * 2.0**-32, (f + 1.0) * 2.0**-32,
                   (f + 1.5) * 2.0**-32], default=(f - 0.5) * 2.0**-32)
    A = np.clip(A, 0.0, 1.0)
    dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a.astype(dt))).cuda()
    dA, df, dc, dw = dev(A, np.float64), dev(f, np.uint32).view(torch.int32), dev(c0, np.uint32).view(torch.int32), dev(word, np.uint32).view(torch.int32)
    dec = torch.zeros(n, dtype=torch.uint8, device='cuda')
    u = torch.zeros(n, dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().svb_debug_decide_lazy(dA.data_ptr(), df.data_ptr(), dc.data_ptr(), dw.data_ptr(), n, stream_id, seed, chain,
                                                 sweep, dec.data_ptr(), u.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert (u.cpu().numpy() == u_ref).all()
    assert (dec.cpu().numpy().astype(bool) == (u_ref < A)).all()


def test_uniform_does_not_depend_on_wide_proposals():
    """Given the four base-K digits, the remainder of word B only takes every K^4-th value, so P(accept | proposal) would
    be quantised in units of K^4 2^-32.  At K = 3 (interval_n = 1) that is 1.9e-8 and the mapping keeps the remainder; for
    K^4 > 256 the leading bits of u must come from elsewhere (the refinement block).  Check both regimes in the oracle's
    statement of the mapping (the kernels are compared with it bit for bit in the GPU tests)."""
    N = 64
    for interval_n, lattice_expected in ((1, True), (2, False), (5, False)):
        K4 = (2 * interval_n + 1) ** 4
        d = P.villain_draws(seed=99, chain=3, sweep=1, N=N, interval_n=interval_n)
        lead = np.floor(d['u'] * 2.0**32).astype(np.int64)                 # the leading 32 bits of the uniform
        digits = np.stack([d['dn_fwd'][0], d['dn_bwd'][0], d['dn_fwd'][1], d['dn_bwd'][1]]).reshape(4, -1) + interval_n
        code = ((digits[0] * (2 * interval_n + 1) + digits[1]) * (2 * interval_n + 1) + digits[2]) * (2 * interval_n + 1) + digits[3]
        # remainder of word B after digits `code`: f = K^4 B - code 2^32, so (f + code 2^32) is a multiple of K^4
        on_lattice = ((lead.reshape(-1) + code * 2**32) % K4 == 0)
        if lattice_expected:
            assert on_lattice.all()
        else:
            assert on_lattice.mean() < 5.0 / K4 + 0.01
