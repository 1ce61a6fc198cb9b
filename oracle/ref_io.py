"""numpy reader/writer for the reference's io/ files (test infrastructure, NOT product code).

Formats are bincode 1.3 default config (little-endian, fixed-width ints, u64 length prefix per
Vec, fields in declaration order) of the Rust types the stage binaries (de)serialise:
  bsk.bin        LweBootstrapKeyOwned<u64>          server_encrypted_aes_decryption.rs:630
  ksk.bin        GlweKeyswitchKeyOwned<u64>         cbs_lib/src/glwe_keyswitch.rs:8-20
  ss_key.bin     GgswCiphertextListOwned<u64>       server_encrypted_aes_decryption.rs:632
  auto_keys.bin  HashMap<usize, AutomorphKeySerializable>   cbs_lib/src/automorphism.rs:237-245
  trans_key.bin  AllRdKeys                          src/data_struct.rs:11-26
  result.bin     LweCiphertextList<Vec<u64>>        server_encrypted_aes_decryption.rs:700-704
  secret_keys/*  LweSecretKey / GlweSecretKey       src/bin/client_key_generation.rs:114-121

This module is independent of the product's C++ reader (csrc/host/bincode_io.cpp) on purpose:
tests cross-check the two.
"""
import os
import struct
import numpy as np

N = 1024
K = 2
GLWE_WORDS = (K + 1) * N
MOD_TAIL = 16 + 8  # CiphertextModulus u128 + scalar_bits u64


class _Cur:
    def __init__(self, buf):
        self.b = memoryview(buf)
        self.o = 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def vec_u64(self):
        n = self.u64()
        a = np.frombuffer(self.b, dtype="<u8", count=n, offset=self.o)
        self.o += 8 * n
        return a

    def modulus(self):
        lo, hi, bits = struct.unpack_from("<QQQ", self.b, self.o)
        self.o += 24
        assert lo == 0 and hi == 0 and bits == 64, "native 2^64 modulus expected"

    def done(self):
        assert self.o == len(self.b), f"trailing bytes: {len(self.b) - self.o}"


def read_bsk(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    glwe_size, poly, base_log, level = c.u64(), c.u64(), c.u64(), c.u64()
    c.modulus()
    c.done()
    assert (glwe_size, poly, base_log, level) == (3, 1024, 23, 1)
    return data.reshape(768, 1, 3, 3, 1024)


def read_ksk(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    in_dim, out_dim, poly, base_log, level = c.u64(), c.u64(), c.u64(), c.u64(), c.u64()
    c.modulus()
    c.done()
    assert (in_dim, out_dim, poly, base_log, level) == (8, 3, 256, 4, 3)
    return data.reshape(8, 3, 4, 256)


def read_ss_key(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    glwe_size, poly, base_log, level = c.u64(), c.u64(), c.u64(), c.u64()
    c.modulus()
    c.done()
    assert (glwe_size, poly, base_log, level) == (3, 1024, 17, 2)
    return data.reshape(2, 2, 3, 3, 1024)


def _twist(n_real):
    j = np.arange(n_real // 2)
    return np.exp(1j * np.pi * j / n_real)


def fourier_to_std_poly(F, order="natural"):
    """Invert tfhe's forward_as_torus for one polynomial (N/2 complex -> N real, units of 2^-64).

    Forward convention (SURVEY.md 8(b) note 1): z_j = (a_j + i a_{j+N/2}) e^{i pi j/N} 2^-64,
    X_m = sum_j z_j e^{-2 pi i j m/(N/2)}; stored order natural in m.
    """
    n = F.shape[-1]
    z = np.fft.ifft(F, axis=-1) * np.conj(_twist(2 * n))
    return np.concatenate([z.real, z.imag], axis=-1)


def std_to_fourier_poly(a_real):
    """forward_as_torus in natural order: a_real already scaled (float64, units of 1)."""
    n = a_real.shape[-1] // 2
    z = (a_real[..., :n] + 1j * a_real[..., n:]) * _twist(2 * n)
    return np.fft.fft(z, axis=-1)


def read_auto_keys(path, split=41):
    """Returns (auto_std[10][2][3][3][1024] u64 with index i <-> kappa=(1024>>i)+1, max_abs_err).

    The file holds Fourier-domain split limbs [in 2][split 2: lo41, hi23][level 3][poly 3][512] c64
    (cbs_lib/src/automorphism.rs:248-254 + cbs_lib/src/fourier_glwe_keyswitch.rs:188-199).  We go
    back to the standard domain and verify that every limb coefficient is an integer in range.
    """
    c = _Cur(open(path, "rb").read())
    count = c.u64()
    assert count == 10
    out = np.zeros((10, 2, 3, 3, N), dtype=np.uint64)
    worst = 0.0
    for _ in range(count):
        kappa = c.u64()
        data = c.vec_u64()
        base_log, level, glwe_dim, poly, auto_k = c.u64(), c.u64(), c.u64(), c.u64(), c.u64()
        assert (base_log, level, glwe_dim, poly) == (13, 3, 2, 1024) and auto_k == kappa
        idx = {(N >> i) + 1: i for i in range(10)}[kappa]
        F = data.view("<f8").reshape(2, 2, 3, 3, 512, 2)
        F = F[..., 0] + 1j * F[..., 1]
        a = fourier_to_std_poly(F) * 2.0 ** 64  # [2][2][3][3][1024]
        r = np.rint(a)
        # lo limb is 41 bits wide: f64 round-off of the stored spectrum leaves ~2^-9 absolute;
        # the 23-bit hi limb must come back as exact integers (SURVEY.md 8(b) note 1).
        worst = max(worst, float(np.abs(a - r)[:, 1].max()))
        assert float(np.abs(a - r)[:, 0].max()) < 0.05, "lo limb not integral: wrong FFT order?"
        lo, hi = r[:, 0], r[:, 1]
        assert lo.min() >= 0 and lo.max() < 2.0 ** split, "lo limb out of range: wrong FFT order?"
        assert hi.min() >= 0 and hi.max() < 2.0 ** (64 - split), "hi limb out of range: wrong FFT order?"
        out[idx] = lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(split))
    c.done()
    assert worst < 1e-4, f"auto key limbs are not integers (err {worst}): unexpected FFT ordering"
    return out, worst


def _glwe_list(c):
    data = c.vec_u64()
    glwe_size, poly = c.u64(), c.u64()
    c.modulus()
    assert (glwe_size, poly) == (3, 1024)
    return data.reshape(-1, GLWE_WORDS)


def _glwe(c):
    data = c.vec_u64()
    poly = c.u64()
    c.modulus()
    assert poly == 1024 and data.size == GLWE_WORDS
    return data


def read_trans_key(path):
    """Returns (k10_9[4][16][2][3072], k8_1[8][4][16][2][3072], k0[16][2][3072]); multiples order
    x9, x11, x13, x14 (src/data_struct.rs:76-81)."""
    c = _Cur(open(path, "rb").read())
    k10_9 = np.zeros((4, 16, 2, GLWE_WORDS), dtype=np.uint64)
    for m in range(4):
        assert c.u64() == 16
        for b in range(16):
            k10_9[m, b] = _glwe_list(c)
    assert c.u64() == 8
    k8_1 = np.zeros((8, 4, 16, 2, GLWE_WORDS), dtype=np.uint64)
    for r in range(8):
        for m in range(4):
            assert c.u64() == 16
            for b in range(16):
                assert c.u64() == 2
                for a in range(2):
                    k8_1[r, m, b, a] = _glwe(c)
    assert c.u64() == 16
    k0 = np.zeros((16, 2, GLWE_WORDS), dtype=np.uint64)
    for b in range(16):
        assert c.u64() == 2
        for a in range(2):
            k0[b, a] = _glwe(c)
    c.done()
    return k10_9, k8_1, k0


def read_lwe_list(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    lwe_size = c.u64()
    c.modulus()
    c.done()
    return data.reshape(-1, lwe_size)


def write_lwe_list(path, arr):
    arr = np.ascontiguousarray(arr, dtype="<u8")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", arr.size))
        f.write(arr.tobytes())
        f.write(struct.pack("<QQQQ", arr.shape[-1], 0, 0, 64))


def read_lwe_sk(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    c.done()
    return data


def read_glwe_sk(path):
    c = _Cur(open(path, "rb").read())
    data = c.vec_u64()
    poly = c.u64()
    c.done()
    assert poly == 1024
    return data


def read_db_hex(path):
    return bytes.fromhex(open(path).read().strip())


def load_server_inputs(io_dir):
    pk = os.path.join(io_dir, "public_keys")
    auto_std, err = read_auto_keys(os.path.join(pk, "auto_keys.bin"))
    return dict(
        bsk=read_bsk(os.path.join(pk, "bsk.bin")),
        ksk=read_ksk(os.path.join(pk, "ksk.bin")),
        ss=read_ss_key(os.path.join(pk, "ss_key.bin")),
        auto_std=auto_std,
        auto_roundtrip_err=err,
        trans_key=read_trans_key(os.path.join(io_dir, "ciphertexts_upload", "trans_key.bin")),
    )


# ---- decryption helpers (client_decrypt_decode*.rs + src/help_fun.rs:24-42) ----
def lwe_phase(lwe, sk):
    """b - <a, s> mod 2^64 for a list of LWE ciphertexts [.., n+1]."""
    lwe = np.asarray(lwe, dtype=np.uint64)
    sk = np.asarray(sk, dtype=np.uint64)
    with np.errstate(over="ignore"):
        dot = (lwe[..., :-1] * sk).sum(axis=-1, dtype=np.uint64)
        return lwe[..., -1] - dot


def decode_bit(phase):
    """help_fun.rs:38-41 with delta = 2^63."""
    phase = np.asarray(phase, dtype=np.uint64)
    with np.errstate(over="ignore"):
        rounding = (phase & np.uint64(1 << 62)) << np.uint64(1)
        return ((phase + rounding) >> np.uint64(63)).astype(np.uint8)


def bit_error(phase, bit):
    """signed distance of the phase from bit*2^63 (float64)."""
    phase = np.asarray(phase, dtype=np.uint64)
    with np.errstate(over="ignore"):
        e = phase - (np.asarray(bit, dtype=np.uint64) << np.uint64(63))
    return e.astype(np.int64).astype(np.float64)


def noise_stats(lwe, sk):
    ph = lwe_phase(lwe, sk)
    bits = decode_bit(ph)
    err = bit_error(ph, bits)
    return bits, float(np.log2(np.sqrt(np.mean(err ** 2)) + 1.0)), float(np.log2(np.abs(err).max() + 1.0))


def bits_to_u16(bits):
    """client_postprocess*.rs:21-28 — 16 bits MSB-first per value."""
    bits = np.asarray(bits).reshape(-1, 16)
    w = (1 << np.arange(15, -1, -1)).astype(np.uint32)
    return (bits * w).sum(axis=1).tolist()
