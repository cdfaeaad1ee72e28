"""
oracle/sample_gen.py — TEST INFRASTRUCTURE: CPU restatement of the on-device obstacle-sample generator
(SURVEY.md §8-f2).  Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module.

What it replaces: simulation/obstacles.py:43-77 of the reference (`generate_obstacle_sample_trajectories`):
    sample_trajectories[:, t, :] = nominal_trajectory[t, :] + np.random.multivariate_normal(0, noise_cov, n_samples)
i.e. per (obstacle, step) N i.i.d. Gaussian samples around the nominal position.  The reference draws them from
numpy's legacy global MT19937 stream with the polar (rejection) Gaussian method: the number of uniforms consumed per
normal is data dependent, so the stream position of halfspace b cannot be computed without replaying all earlier
draws — it cannot be reproduced by independent GPU threads.  PARITY UNPINNED against the reference's random stream
(by construction); the DISTRIBUTION is the reference's (mean = nominal position, covariance = noise_cov), and the
generator itself is pinned bit for bit between this file and the CUDA kernel:

  * counter-based Philox4x32-10 (Salmon et al., SC'11; known-answer vectors in tests/test_sample_gen.py),
    key = (seed & 0xffffffff, seed >> 32), counter = (pair index j, halfspace index b low, b high, 0x44524356);
    one call gives r0..r3 -> samples 2j (r0, r1) and 2j+1 (r2, r3) of halfspace b;
  * u1 = ((r >> 8) + 1) * 2^-24 in (0, 1],  u2 = (r' >> 8) * 2^-24 in [0, 1)      (exact in fp32);
  * Box-Muller in fp32 with explicitly ordered +,-,*,/ and sqrt only (all IEEE round-to-nearest, no FMA), so that numpy
    float32 arithmetic and the CUDA kernel (__fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn) agree bit for bit:
        rad = sqrt(-2 log(u1)),  (z0, z1) = rad * (cos(2 pi u2), sin(2 pi u2));
    log(): fdlibm-style reduction to [sqrt(1/2), sqrt(2)) + degree-4 polynomial in s^2, s = f/(2+f);
    sin/cos(2 pi u2): exact quadrant split t = 4 u2, q = floor(t), r = t - q, reflection at r > 1/2, Taylor kernels on
    [0, pi/4] (|error| < 3e-8);
  * sample = (mx + l00 z0,  my + (l10 z0 + l11 z1)) with L = chol(noise_cov) and the mean rounded to fp32 first.
"""
import numpy as np

F = np.float32
U32 = np.uint32
U64 = np.uint64

PHILOX_M0 = U64(0xD2511F53)
PHILOX_M1 = U64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
STREAM_TAG = 0x44524356   # 'DRCV'

# fp32 constants (identical literals in csrc/sample_gen.cuh)
LN2_HI = F(float.fromhex("0x1.62e300p-1"))     # 6.9313812256e-01
LN2_LO = F(float.fromhex("0x1.2fefa2p-17"))    # 9.0580006145e-06
LG1 = F(float.fromhex("0x1.555554p-1"))        # 0.66666662693
LG2 = F(float.fromhex("0x1.999c26p-2"))        # 0.40000972152
LG3 = F(float.fromhex("0x1.23d3dcp-2"))        # 0.28498786688
LG4 = F(float.fromhex("0x1.f13c4cp-3"))        # 0.24279078841
HALF_PI = F(float.fromhex("0x1.921fb6p+0"))    # pi/2 rounded to fp32
S1 = F(-1.0 / 6.0)
S2 = F(1.0 / 120.0)
S3 = F(-1.0 / 5040.0)
S4 = F(1.0 / 362880.0)
C1 = F(1.0 / 24.0)
C2 = F(-1.0 / 720.0)
C3 = F(1.0 / 40320.0)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (counters) with scalar keys; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=U32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = U64(0xFFFFFFFF)
    for _ in range(10):
        p0 = c0.astype(U64) * PHILOX_M0
        p1 = c2.astype(U64) * PHILOX_M1
        hi0, lo0 = (p0 >> U64(32)).astype(U32), (p0 & mask).astype(U32)
        hi1, lo1 = (p1 >> U64(32)).astype(U32), (p1 & mask).astype(U32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ U32(k0), lo1, hi0 ^ c3 ^ U32(k1), lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def log_f32(x):
    """fp32 log for x in [2^-24, 1] (normal floats), explicit op order; mirrors csrc/sample_gen.cuh::gen_log."""
    x = np.asarray(x, dtype=F)
    ix = x.view(np.uint32).astype(np.int64)
    ix = ix + (0x3F800000 - 0x3F3504F3)
    k = (ix >> 23) - 0x7F
    ix = (ix & 0x007FFFFF) + 0x3F3504F3
    m = ix.astype(np.uint32).view(F)
    f = m - F(1.0)
    s = f / (F(2.0) + f)
    z = s * s
    w = z * z
    t1 = w * (LG2 + w * LG4)
    t2 = z * (LG1 + w * LG3)
    r = t2 + t1
    hfsq = (F(0.5) * f) * f
    dk = k.astype(F)
    return (((s * (hfsq + r)) + (dk * LN2_LO)) - hfsq + f) + (dk * LN2_HI)


def sincos_2pi_f32(u):
    """(cos(2 pi u), sin(2 pi u)) for u in [0, 1), fp32, explicit op order; mirrors gen_sincos."""
    u = np.asarray(u, dtype=F)
    t = F(4.0) * u
    qf = np.floor(t)
    r = t - qf
    q = qf.astype(np.int32)
    flip = r > F(0.5)
    rr = np.where(flip, F(1.0) - r, r).astype(F)
    x = rr * HALF_PI
    z = x * x
    sp = x + (x * z) * (S1 + z * (S2 + z * (S3 + z * S4)))
    cp = (F(1.0) - F(0.5) * z) + (z * z) * (C1 + z * (C2 + z * C3))
    s_ = np.where(flip, cp, sp).astype(F)
    c_ = np.where(flip, sp, cp).astype(F)
    cos_o = np.where(q == 0, c_, np.where(q == 1, -s_, np.where(q == 2, -c_, s_))).astype(F)
    sin_o = np.where(q == 0, s_, np.where(q == 1, c_, np.where(q == 2, -s_, -c_))).astype(F)
    return cos_o, sin_o


def normals_from_bits(ra, rb):
    """One Box-Muller pair (z0, z1) in fp32 from two uint32 words."""
    u1 = (((ra >> U32(8)).astype(np.int64) + 1).astype(F)) * F(2.0 ** -24)
    u2 = (rb >> U32(8)).astype(F) * F(2.0 ** -24)
    rad = np.sqrt(F(-2.0) * log_f32(u1))
    c, s = sincos_2pi_f32(u2)
    return rad * c, rad * s


def cholesky2(cov):
    """Lower Cholesky factor (l00, l10, l11) of a 2x2 covariance, fp64 (the caller passes it to the kernel as fp64)."""
    cov = np.asarray(cov, dtype=np.float64)
    l00 = np.sqrt(cov[..., 0, 0])
    l10 = np.where(l00 > 0, cov[..., 1, 0] / np.where(l00 > 0, l00, 1.0), 0.0)
    l11 = np.sqrt(np.maximum(cov[..., 1, 1] - l10 * l10, 0.0))
    return np.stack([l00, l10, l11], axis=-1)


def generate(mean, chol, n_samples, seed, index_offset=0):
    """
    Samples of B halfspaces: mean [B,2] fp64, chol [B,3] fp64 (l00, l10, l11), -> float32 [B, n_samples, 2].
    Halfspace b uses Philox counter word (index_offset + b); identical to the kernel's generate mode.
    """
    mean = np.atleast_2d(np.asarray(mean, dtype=np.float64)).astype(F)
    chol = np.atleast_2d(np.asarray(chol, dtype=np.float64)).astype(F)
    B = mean.shape[0]
    n_pairs = (n_samples + 1) // 2
    j = np.arange(n_pairs, dtype=np.uint64)[None, :]
    b = (np.arange(B, dtype=np.uint64) + np.uint64(index_offset))[:, None]
    r0, r1, r2, r3 = philox4x32_10((j & U64(0xFFFFFFFF)).astype(U32), (b & U64(0xFFFFFFFF)).astype(U32),
                                   (b >> U64(32)).astype(U32), U32(STREAM_TAG),
                                   int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    out = np.empty((B, 2 * n_pairs, 2), dtype=F)
    for e, (ra, rb) in enumerate(((r0, r1), (r2, r3))):
        z0, z1 = normals_from_bits(ra, rb)
        out[:, e::2, 0] = mean[:, 0:1] + (chol[:, 0:1] * z0)
        out[:, e::2, 1] = mean[:, 1:2] + ((chol[:, 1:2] * z0) + (chol[:, 2:3] * z1))
    return np.ascontiguousarray(out[:, :n_samples, :])
