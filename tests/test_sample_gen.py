"""
CPU tests of oracle/sample_gen.py — the restatement of the on-device sample generator:
Philox4x32-10 known-answer vectors (Random123 kat_vectors), accuracy of the fp32 log / sincos kernels, the distribution
of the generated samples (the reference's: nominal + N(0, noise_cov), simulation/obstacles.py:68-75), and the
counter layout (sharding by index_offset reproduces the unsharded stream).
"""
import numpy as np

from oracle import sample_gen as sg


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    cases = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in cases:
        got = sg.philox4x32_10(*(np.array([c], dtype=np.uint32) for c in ctr), key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want, (ctr, key, [hex(int(g[0])) for g in got])


def test_log_and_sincos_accuracy():
    rng = np.random.RandomState(0)
    u = np.concatenate([(rng.randint(0, 2 ** 24, size=20000) + 1) * 2.0 ** -24, [2.0 ** -24, 1.0, 0.5, 1 - 2.0 ** -24]])
    got = sg.log_f32(u.astype(np.float32)).astype(np.float64)
    assert got.dtype == np.float64 and np.all(got <= 0.0)
    assert np.abs(got - np.log(u)).max() <= 2e-7 * np.maximum(1.0, np.abs(np.log(u))).max()
    v = np.concatenate([rng.randint(0, 2 ** 24, size=20000) * 2.0 ** -24, [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -24]])
    c, s = sg.sincos_2pi_f32(v.astype(np.float32))
    assert c.dtype == np.float32 and s.dtype == np.float32
    assert np.abs(c - np.cos(2 * np.pi * v)).max() < 3e-7 and np.abs(s - np.sin(2 * np.pi * v)).max() < 3e-7


def test_distribution_matches_reference_model():
    """mean = nominal position, covariance = noise_cov (obstacles.py:68-75), standard-normal marginals."""
    cov = np.array([[0.01, 0.004], [0.004, 0.02]])
    mean = np.array([[3.0, -1.0], [0.5, 0.25]])
    s = sg.generate(mean, sg.cholesky2(cov), 200000, seed=42).astype(np.float64)
    assert s.shape == (2, 200000, 2)
    for b in range(2):
        assert np.abs(s[b].mean(axis=0) - mean[b]).max() < 1e-3
        assert np.abs(np.cov(s[b].T) - cov).max() < 3e-4
    z = (s[0, :, 0] - 3.0) / 0.1
    assert abs(np.mean(z ** 3)) < 0.03 and abs(np.mean(z ** 4) - 3.0) < 0.06
    assert abs(np.mean(np.abs(z) > 3.0) - 0.0027) < 5e-4
    # zero covariance (the reference's t = 0 row, obstacles.py:63): every sample is the nominal position
    s0 = sg.generate(mean[:1], np.zeros((1, 3)), 7, seed=1)
    assert np.array_equal(s0, np.broadcast_to(mean[:1].astype(np.float32)[:, None, :], (1, 7, 2)))


def test_counter_layout_and_determinism():
    chol = sg.cholesky2(np.diag([0.01, 0.01]))
    mean = np.arange(12, dtype=np.float64).reshape(6, 2)
    a = sg.generate(mean, chol, 1001, seed=7)
    assert np.array_equal(a, sg.generate(mean, chol, 1001, seed=7))
    assert not np.array_equal(a, sg.generate(mean, chol, 1001, seed=8))
    # a shard of the batch reproduces its rows; a shorter N is a prefix of a longer one
    assert np.array_equal(a[4:], sg.generate(mean[4:], chol, 1001, seed=7, index_offset=4))
    assert np.array_equal(a[:, :500], sg.generate(mean, chol, 500, seed=7))
    # halfspaces with the same mean still get different draws
    same = sg.generate(np.zeros((2, 2)), chol, 64, seed=7)
    assert not np.array_equal(same[0], same[1])
