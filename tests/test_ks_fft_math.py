"""The index arithmetic of the keyswitch's 128-point transform (csrc/cbs_kernels.cu, ks_fwd_a / ks_fwd_bc / ks_inv_bc /
ks_inv_a), restated in numpy: pass A = radix 8 over the stride-16 points of thread j with twiddle W128^(j k), pass BC = radix
16 over the contiguous block of thread k, positions XOR-swizzled inside each block of 16.  Checks (1) that position
ks_pos(k, g) holds frequency k + 8 g of the plain DFT, (2) that the mirrored inverse returns 128 x the input, (3) that both
passes touch shared memory without bank conflicts (16-byte accesses, 8 lanes per 128-byte wavefront)."""
import numpy as np


def ks_pos(k, j):
    return 16 * k + (j ^ k)


def dft8(v, inv=False):
    k = np.arange(8)
    return np.exp((2j if inv else -2j) * np.pi * np.outer(k, k) / 8) @ v


def w(n, e):
    return np.exp(-2j * np.pi * e / n)


def forward(x):
    F = np.zeros(128, complex)
    for j in range(16):  # pass A, thread j
        v = dft8(np.array([x[j + 16 * m] for m in range(8)]))
        for k in range(8):
            F[ks_pos(k, j)] = v[k] * w(128, j * k)
    for k in range(8):  # pass BC, thread k
        y = np.array([F[ks_pos(k, jj)] for jj in range(16)])
        e, o = dft8(y[0::2]), dft8(y[1::2])
        for g in range(8):
            t = o[g] * w(16, g)
            F[ks_pos(k, g)], F[ks_pos(k, g + 8)] = e[g] + t, e[g] - t
    return F


def inverse(F):
    F = F.copy()
    for k in range(8):
        z = np.array([F[ks_pos(k, g)] for g in range(16)])
        e = dft8(z[:8] + z[8:], True)
        o = dft8((z[:8] - z[8:]) * np.conj(w(16, np.arange(8))), True)
        for a in range(8):
            F[ks_pos(k, 2 * a)], F[ks_pos(k, 2 * a + 1)] = e[a], o[a]
    x = np.zeros(128, complex)
    for j in range(16):
        v = dft8(np.array([F[ks_pos(k, j)] * np.conj(w(128, j * k)) for k in range(8)]), True)
        for m in range(8):
            x[j + 16 * m] = v[m]
    return x


def test_positions_hold_the_plain_dft_and_the_inverse_mirrors_it():
    rng = np.random.default_rng(7)
    x = rng.normal(size=128) + 1j * rng.normal(size=128)
    F = forward(x)
    X = np.fft.fft(x)
    assert max(abs(F[ks_pos(k, g)] - X[k + 8 * g]) for k in range(8) for g in range(16)) < 1e-11
    assert np.abs(inverse(F) / 128 - x).max() < 1e-12


def test_negacyclic_product_through_the_transform():
    # the keyswitch multiplies folded + twisted polynomials of the ring Z[X]/(X^256 + 1) pointwise in this position order
    rng = np.random.default_rng(8)
    a = rng.integers(-8, 9, 256).astype(float)
    b = rng.integers(-1000, 1000, 256).astype(float)
    twist = np.exp(1j * np.pi * np.arange(128) / 256)
    fold = lambda p: (p[:128] + 1j * p[128:]) * twist
    z = inverse(forward(fold(a)) * forward(fold(b))) / 128 * np.conj(twist)
    got = np.concatenate([z.real, z.imag])
    want = np.zeros(256)
    for i in range(256):
        for j in range(256):
            s = i + j
            want[s % 256] += (a[i] * b[j]) * (1 if s < 256 else -1)
    assert np.abs(got - want).max() < 1e-6


def test_both_passes_are_bank_conflict_free():
    # a 16-byte access is served 8 lanes at a time; the 8 lanes must hit 8 different 16-byte slots modulo 128 bytes
    for k in range(8):  # pass A stores (and inverse A loads): lanes j = 0..15 at fixed register k
        for half in range(2):
            slots = {ks_pos(k, j) % 8 for j in range(8 * half, 8 * half + 8)}
            assert len(slots) == 8
    for i in range(16):  # pass BC loads / stores: lanes k = 0..7 at fixed register index i
        assert len({ks_pos(k, i) % 8 for k in range(8)}) == 8
