"""Cleartext AES-128 (test infrastructure, NOT product code).

Plain FIPS-197 AES-128 block encrypt/decrypt + CTR keystream, used by the
tests, the oracle pin script and bench.py's verifier to produce `db.hex` and the
expected plaintexts the way the reference harness does with pyaes
(reference: harness/aes_keygen_and_encrypt.py:33-55, harness/cleartext_impl.py:33-50).
pyaes is not installed in this image, so this is a from-scratch restatement of
the standard cipher; it is checked against the FIPS-197 Appendix C.1 vector in
tests/test_oracle_cpu.py.
"""
import hashlib
import struct

def _xtime(a):
    a <<= 1
    return (a ^ 0x11B) & 0xFF if a & 0x100 else a

def _gmul(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        a = _xtime(a)
        b >>= 1
    return r

def _build_sbox():
    # multiplicative inverse in GF(2^8) followed by the affine map
    inv = [0] * 256
    for x in range(1, 256):
        for y in range(1, 256):
            if _gmul(x, y) == 1:
                inv[x] = y
                break
    sbox = [0] * 256
    for x in range(256):
        b = inv[x]
        r = 0
        for i in range(8):
            bit = ((b >> i) ^ (b >> ((i + 4) % 8)) ^ (b >> ((i + 5) % 8)) ^
                   (b >> ((i + 6) % 8)) ^ (b >> ((i + 7) % 8)) ^ (0x63 >> i)) & 1
            r |= bit << i
        sbox[x] = r
    return sbox

SBOX = _build_sbox()
INV_SBOX = [0] * 256
for _i, _v in enumerate(SBOX):
    INV_SBOX[_v] = _i
RCON = [0x00, 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36]

def gmul(a, b):
    return _gmul(a, b)

def expand_key(key):
    """Return 11 round keys, each a list of 16 bytes (byte index = 4*col+row)."""
    w = [list(key[4 * i:4 * i + 4]) for i in range(4)]
    for i in range(4, 44):
        t = list(w[i - 1])
        if i % 4 == 0:
            t = t[1:] + t[:1]
            t = [SBOX[b] for b in t]
            t[0] ^= RCON[i // 4]
        w.append([a ^ b for a, b in zip(w[i - 4], t)])
    return [sum((w[4 * r + c] for c in range(4)), []) for r in range(11)]

def _shift_rows(s, inv=False):
    o = [0] * 16
    for c in range(4):
        for r in range(4):
            if inv:
                o[4 * c + r] = s[4 * ((c - r) % 4) + r]
            else:
                o[4 * c + r] = s[4 * ((c + r) % 4) + r]
    return o

def _mix_columns(s, inv=False):
    m = [14, 11, 13, 9] if inv else [2, 3, 1, 1]
    o = [0] * 16
    for c in range(4):
        col = s[4 * c:4 * c + 4]
        for r in range(4):
            o[4 * c + r] = (_gmul(col[r], m[0]) ^ _gmul(col[(r + 1) % 4], m[1]) ^
                            _gmul(col[(r + 2) % 4], m[2]) ^ _gmul(col[(r + 3) % 4], m[3]))
    return o

def encrypt_block(key, pt):
    rk = expand_key(key)
    s = [a ^ b for a, b in zip(pt, rk[0])]
    for r in range(1, 10):
        s = [SBOX[b] for b in s]
        s = _shift_rows(s)
        s = _mix_columns(s)
        s = [a ^ b for a, b in zip(s, rk[r])]
    s = [SBOX[b] for b in s]
    s = _shift_rows(s)
    return bytes(a ^ b for a, b in zip(s, rk[10]))

def decrypt_block(key, ct):
    rk = expand_key(key)
    s = [a ^ b for a, b in zip(ct, rk[10])]
    for r in range(9, 0, -1):
        s = _shift_rows(s, inv=True)
        s = [INV_SBOX[b] for b in s]
        s = [a ^ b for a, b in zip(s, rk[r])]
        s = _mix_columns(s, inv=True)
    s = _shift_rows(s, inv=True)
    s = [INV_SBOX[b] for b in s]
    return bytes(a ^ b for a, b in zip(s, rk[0]))

def ecb_encrypt(key, data):
    return b"".join(encrypt_block(key, data[i:i + 16]) for i in range(0, len(data), 16))

def ecb_decrypt(key, data):
    return b"".join(decrypt_block(key, data[i:i + 16]) for i in range(0, len(data), 16))

def ctr_keystream_blocks(key, iv, nblocks):
    """pyaes.Counter semantics: 128-bit big-endian counter, +1 per block."""
    ctr = int.from_bytes(iv, "big")
    out = []
    for _ in range(nblocks):
        out.append(encrypt_block(key, (ctr % (1 << 128)).to_bytes(16, "big")))
        ctr += 1
    return out

def ctr_crypt(key, iv, data):
    ks = b"".join(ctr_keystream_blocks(key, iv, (len(data) + 15) // 16))
    return bytes(a ^ b for a, b in zip(data, ks))

def harness_aes_key(seed=None):
    """harness/aes_keygen_and_encrypt.py:33 — sha256(str(seed))[:16]."""
    return hashlib.sha256(str(seed).encode()).digest()[:16]

def harness_iv(seed=None):
    return hashlib.sha256(b"iv" + str(seed).encode()).digest()[:16]

def pack_u16_be(values):
    return struct.pack(">" + "H" * len(values), *values)

def unpack_u16_be(data):
    return list(struct.unpack(">" + "H" * (len(data) // 2), data))
