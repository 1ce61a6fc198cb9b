"""pyaes stand-in for the reference's Python harness (test infrastructure, not product code).

harness/aes_keygen_and_encrypt.py:12,45-55 and harness/cleartext_impl.py:18,39-52 import `pyaes`
(requirements.txt), which is not installed in this image and cannot be fetched (no network).  This module
provides exactly the three names they use, with pyaes' call conventions, on top of the `cryptography`
package (OpenSSL) that the image does ship - so the harness's cleartext expectation comes from an AES
implementation that is independent of this repository's.  tests/harness_run.py puts this directory on
PYTHONPATH for the harness subprocesses only.

    AES(key).encrypt(block16) / .decrypt(block16)   -> list of 16 ints   (one ECB block, like pyaes)
    Counter(initial_value)                          -> 128-bit big-endian counter, +1 per block
    AESModeOfOperationCTR(key, counter=Counter(n)).encrypt(data) / .decrypt(data) -> bytes
"""
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes

__all__ = ["AES", "Counter", "AESModeOfOperationCTR"]


def _bytes(x):
    return bytes(bytearray(x))


class AES(object):
    def __init__(self, key):
        if len(key) not in (16, 24, 32):
            raise ValueError("Invalid key size")
        self._key = _bytes(key)

    def encrypt(self, plaintext):
        if len(plaintext) != 16:
            raise ValueError("wrong block length")
        enc = Cipher(algorithms.AES(self._key), modes.ECB()).encryptor()
        return list(enc.update(_bytes(plaintext)) + enc.finalize())

    def decrypt(self, ciphertext):
        if len(ciphertext) != 16:
            raise ValueError("wrong block length")
        dec = Cipher(algorithms.AES(self._key), modes.ECB()).decryptor()
        return list(dec.update(_bytes(ciphertext)) + dec.finalize())


class Counter(object):
    def __init__(self, initial_value=1):
        self._value = int(initial_value) % (1 << 128)

    value = property(lambda self: self._value.to_bytes(16, "big"))

    def increment(self):
        self._value = (self._value + 1) % (1 << 128)


class AESModeOfOperationCTR(object):
    name = "Counter (CTR)"

    def __init__(self, key, counter=None):
        self._key = _bytes(key)
        self._counter = counter if counter is not None else Counter()
        self._stream = Cipher(algorithms.AES(self._key), modes.CTR(self._counter.value)).encryptor()

    def encrypt(self, plaintext):
        return self._stream.update(_bytes(plaintext))

    decrypt = encrypt
