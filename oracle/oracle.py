"""ctypes binding of liboracle.so (test infrastructure, NOT product code — see cbs_oracle.h)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LWE_N = 768
N = 1024
K = 2
BIG = K * N
GLWE_WORDS = (K + 1) * N
GGSW_WORDS = 7 * 3 * GLWE_WORDS

_u64p = ctypes.POINTER(ctypes.c_uint64)
_f64p = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_keys_create.restype = ctypes.c_void_p
        _LIB.orc_modswitch.restype = ctypes.c_uint64
        _LIB.orc_modswitch.argtypes = [ctypes.c_uint64]
    return _LIB


def _u(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Keys:
    def __init__(self, bsk, ksk, auto_std, ss):
        self._k = ctypes.c_void_p(lib().orc_keys_create(_u(_c(bsk)), _u(_c(ksk)), _u(_c(auto_std)), _u(_c(ss))))

    def __del__(self):
        if getattr(self, "_k", None):
            lib().orc_keys_destroy(self._k)
            self._k = None


def modswitch(x):
    return int(lib().orc_modswitch(int(x)))


def lwe_keyswitch(keys, lwe_in):
    lwe_in = _c(lwe_in).reshape(-1, BIG + 1)
    out = np.zeros((lwe_in.shape[0], LWE_N + 1), dtype=np.uint64)
    for i in range(lwe_in.shape[0]):
        lib().orc_lwe_keyswitch(keys._k, _u(lwe_in[i]), _u(out[i]))
    return out


def blind_rotate(keys, lwe):
    lwe = _c(lwe).reshape(-1, LWE_N + 1)
    out = np.zeros((lwe.shape[0], GLWE_WORDS), dtype=np.uint64)
    for i in range(lwe.shape[0]):
        lib().orc_blind_rotate(keys._k, _u(lwe[i]), _u(out[i]))
    return out


def glev_from_acc(acc):
    acc = _c(acc).reshape(-1, GLWE_WORDS)
    out = np.zeros((acc.shape[0], 7, GLWE_WORDS), dtype=np.uint64)
    for i in range(acc.shape[0]):
        lib().orc_glev_from_acc(_u(acc[i]), _u(out[i]))
    return out


def trace(keys, glwe):
    g = _c(glwe).reshape(-1, GLWE_WORDS).copy()
    for i in range(g.shape[0]):
        lib().orc_trace(keys._k, _u(g[i]))
    return g


def auto_step(keys, i, glwe):
    g = _c(glwe).reshape(GLWE_WORDS)
    out = np.zeros(GLWE_WORDS, dtype=np.uint64)
    lib().orc_auto_step(keys._k, ctypes.c_int(i), _u(g), _u(out))
    return out


def scheme_switch(keys, glev):
    glev = _c(glev).reshape(-1, 7, GLWE_WORDS)
    out = np.zeros((glev.shape[0], GGSW_WORDS), dtype=np.uint64)
    for i in range(glev.shape[0]):
        lib().orc_scheme_switch(keys._k, _u(glev[i]), _u(out[i]))
    return out


def circuit_bootstrap(keys, lwe):
    lwe = _c(lwe).reshape(-1, LWE_N + 1)
    out = np.zeros((lwe.shape[0], GGSW_WORDS), dtype=np.uint64)
    for i in range(lwe.shape[0]):
        lib().orc_circuit_bootstrap(keys._k, _u(lwe[i]), _u(out[i]))
    return out


def ggsw_to_fourier(ggsw_std):
    g = _c(ggsw_std).reshape(-1, GGSW_WORDS)
    out = np.zeros((g.shape[0], GGSW_WORDS), dtype=np.float64)
    for i in range(g.shape[0]):
        lib().orc_ggsw_to_fourier(out[i].ctypes.data_as(_f64p), _u(g[i]), ctypes.c_int(K), ctypes.c_int(N), ctypes.c_int(7))
    return out


def lut8_eval(ggsw_f8, lut2):
    f = np.ascontiguousarray(ggsw_f8, dtype=np.float64).reshape(8, GGSW_WORDS)
    lut2 = _c(lut2).reshape(2, GLWE_WORDS)
    out = np.zeros((8, BIG + 1), dtype=np.uint64)
    lib().orc_lut8_eval(f.ctypes.data_as(_f64p), _u(lut2), _u(out))
    return out


def known_rotate(ct16, luts):
    ct = np.frombuffer(bytes(ct16), dtype=np.uint8).copy()
    luts = _c(luts).reshape(16, 2, GLWE_WORDS)
    out = np.zeros((128, BIG + 1), dtype=np.uint64)
    lib().orc_known_rotate(ct.ctypes.data_as(_u8p), _u(luts), _u(out))
    return out


def inv_mix_columns_precomp(t9, t11, t13, t14):
    st = np.zeros((128, BIG + 1), dtype=np.uint64)
    lib().orc_inv_mix_columns_precomp(_u(st), _u(_c(t9)), _u(_c(t11)), _u(_c(t13)), _u(_c(t14)))
    return st


def inv_shift_rows(st):
    st = _c(st).copy()
    lib().orc_inv_shift_rows(_u(st))
    return st


def aes128_transcipher(keys, ct_bytes, k10_9, k8_1, k0):
    ct = np.frombuffer(bytes(ct_bytes), dtype=np.uint8).copy()
    nblocks = ct.size // 16
    out = np.zeros((nblocks, 128, BIG + 1), dtype=np.uint64)
    lib().orc_aes128_transcipher(keys._k, ct.ctypes.data_as(_u8p), ctypes.c_int(nblocks), _u(_c(k10_9)), _u(_c(k8_1)),
                                 _u(_c(k0)), _u(out))
    return out


def aes128_ctr_transcipher(keys, ct_bytes, iv, kf_first, kf_mid, kf_last):
    ct = np.frombuffer(bytes(ct_bytes), dtype=np.uint8).copy()
    ivb = np.frombuffer(bytes(iv), dtype=np.uint8).copy()
    nblocks = ct.size // 16
    out = np.zeros((nblocks, 128, BIG + 1), dtype=np.uint64)
    lib().orc_aes128_ctr_transcipher(keys._k, ct.ctypes.data_as(_u8p), ctypes.c_int(nblocks), ivb.ctypes.data_as(_u8p),
                                     _u(_c(kf_first)), _u(_c(kf_mid)), _u(_c(kf_last)), _u(out))
    return out


def max_of_two(ggsw_a, ggsw_b, lwe_a, lwe_b, reset_e=False):
    out = np.zeros((16, BIG + 1), dtype=np.uint64)
    lib().orc_max_of_two(_u(_c(ggsw_a)), _u(_c(ggsw_b)), _u(_c(lwe_a)), _u(_c(lwe_b)), _u(out), ctypes.c_int(int(reset_e)))
    return out


def max_u16(keys, lwe_in):
    lwe_in = _c(lwe_in).reshape(-1, BIG + 1)
    nvals = lwe_in.shape[0] // 16
    out = np.zeros((16, BIG + 1), dtype=np.uint64)
    lib().orc_max_u16(keys._k, _u(lwe_in), ctypes.c_int(nvals), _u(out))
    return out


def num_threads():
    return int(lib().orc_num_threads())
