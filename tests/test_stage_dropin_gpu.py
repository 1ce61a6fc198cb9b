"""Drop-in test of the stage executables: the REFERENCE's own prebuilt client binaries (oracle/_ref,
unmodified) generate the keys and decrypt the results; OUR server_encrypted_aes_decryption and
server_encrypted_compute run in between, over the same io/ + datasets/ file contract that
harness/run_submission.py drives."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "temp_fhe_transciphering_b200", "bin")


def _run(exe, cwd, *args):
    subprocess.run([exe, "0", *args], cwd=cwd, check=True, stdout=subprocess.DEVNULL, timeout=900)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "client_key_generation")), reason="oracle/_ref not built")
@pytest.mark.parametrize("nvals", [8, 24])
def test_reference_clients_around_our_servers(tmp_path, nvals):
    import aes_clear
    rng = np.random.default_rng(nvals)
    vals = rng.integers(0, 65536, nvals).tolist()
    key = aes_clear.harness_aes_key(None)
    ct = aes_clear.ecb_encrypt(key, aes_clear.pack_u16_be(vals))
    d = tmp_path
    os.makedirs(d / "datasets" / "toy")
    (d / "datasets" / "toy" / "aes_key.hex").write_text(key.hex())
    (d / "datasets" / "toy" / "db.hex").write_text(ct.hex())
    _run(os.path.join(REF, "client_key_generation"), d)          # reference, unseeded
    _run(os.path.join(REF, "client_encode_encrypt"), d)          # reference
    _run(os.path.join(BIN, "server_encrypted_aes_decryption"), d)  # ours (all blocks, all visible GPUs)
    _run(os.path.join(BIN, "server_encrypted_compute"), d)         # ours
    _run(os.path.join(REF, "client_decrypt_decode_aes_decryption"), d)
    _run(os.path.join(REF, "client_postprocess_aes_decryption"), d)
    _run(os.path.join(REF, "client_decrypt_decode"), d)
    _run(os.path.join(REF, "client_postprocess"), d)
    got = [int(x) for x in (d / "io" / "toy" / "result_aes.txt").read_text().split()]
    got_max = [int(x) for x in (d / "io" / "toy" / "result.txt").read_text().split()]
    assert got == vals
    assert got_max == [max(vals)]
    # same bytes on disk as the reference writes: LweCiphertextList of 128 x blocks / 16 ciphertexts
    assert os.path.getsize(d / "io" / "toy" / "ciphertext_aes_download" / "result.bin") == 8 + nvals * 16 * 2049 * 8 + 32
    assert os.path.getsize(d / "io" / "toy" / "ciphertexts_download" / "result.bin") == 262_312


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "client_key_generation")), reason="oracle/_ref not built")
def test_small_instance_ctr_mode(tmp_path):
    """Harness size 1 ("small": 64 u16 = 8 blocks, AES-CTR; harness/aes_keygen_and_encrypt.py:49-55):
    reference key generation and decryption clients, OUR client_encode_encrypt (forward keys; the
    reference's only emits ECB-decryption keys) and OUR two servers."""
    import aes_clear
    rng = np.random.default_rng(64)
    vals = rng.integers(0, 65536, 64).tolist()
    key, iv = aes_clear.harness_aes_key(None), aes_clear.harness_iv(None)
    ct = aes_clear.ctr_crypt(key, iv, aes_clear.pack_u16_be(vals))
    d = tmp_path
    os.makedirs(d / "datasets" / "small")
    (d / "datasets" / "small" / "aes_key.hex").write_text(key.hex())
    (d / "datasets" / "small" / "aes_iv.hex").write_text(iv.hex())
    (d / "datasets" / "small" / "db.hex").write_text(ct.hex())

    def run(exe, *args):
        subprocess.run([exe, "1", *args], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=900)

    run(os.path.join(REF, "client_key_generation"))
    run(os.path.join(BIN, "client_encode_encrypt"))
    run(os.path.join(BIN, "server_encrypted_aes_decryption"))
    run(os.path.join(BIN, "server_encrypted_compute"))
    run(os.path.join(REF, "client_decrypt_decode_aes_decryption"))
    run(os.path.join(REF, "client_postprocess_aes_decryption"))
    run(os.path.join(REF, "client_decrypt_decode"))
    run(os.path.join(REF, "client_postprocess"))
    got = [int(x) for x in (d / "io" / "small" / "result_aes.txt").read_text().split()]
    got_max = [int(x) for x in (d / "io" / "small" / "result.txt").read_text().split()]
    assert got == vals
    assert got_max == [max(vals)]
    # the other mini-workload on the same transciphered values (second argument = harness --mini_workload 1): 32 pairs, sharded
    # over the visible GPUs when there are several, combined with cbs_sum_u16
    run(os.path.join(BIN, "server_encrypted_compute"), "1")
    run(os.path.join(REF, "client_decrypt_decode"))
    run(os.path.join(REF, "client_postprocess"))
    want = sum((x * y) % 65536 for x, y in zip(vals[:32], vals[32:])) % 65536
    assert [int(x) for x in (d / "io" / "small" / "result.txt").read_text().split()] == [want]


def test_all_ten_stages_ours(tmp_path):
    """The whole stage sequence of harness/run_submission.py:69-115 with OUR executables only (seeded
    client stages in C++, GPU server stages), toy instance / ECB."""
    import aes_clear
    rng = np.random.default_rng(77)
    vals = rng.integers(0, 65536, 8).tolist()
    key = aes_clear.harness_aes_key(None)
    d = tmp_path
    os.makedirs(d / "datasets" / "toy")
    (d / "datasets" / "toy" / "aes_key.hex").write_text(key.hex())
    (d / "datasets" / "toy" / "db.hex").write_text(aes_clear.ecb_encrypt(key, aes_clear.pack_u16_be(vals)).hex())
    for exe, extra in (("client_preprocess", []), ("client_key_generation", ["4711"]), ("client_encode_encrypt", ["4712"]),
                       ("server_preprocess_dataset", []), ("server_encrypted_aes_decryption", []),
                       ("server_encrypted_compute", []), ("client_decrypt_decode_aes_decryption", []),
                       ("client_postprocess_aes_decryption", []), ("client_decrypt_decode", []), ("client_postprocess", [])):
        subprocess.run([os.path.join(BIN, exe), "0", *extra], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=900)
    assert [int(x) for x in (d / "io" / "toy" / "result_aes.txt").read_text().split()] == vals
    assert [int(x) for x in (d / "io" / "toy" / "result.txt").read_text().split()] == [max(vals)]
    # mini-workload #2 (harness --mini_workload 1): inner product of the two halves mod 2^16
    env = dict(os.environ, CBS_MINI_WORKLOAD="1")
    for exe in ("server_encrypted_compute", "client_decrypt_decode", "client_postprocess"):
        subprocess.run([os.path.join(BIN, exe), "0"], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=900, env=env)
    want = sum((x * y) % 65536 for x, y in zip(vals[:4], vals[4:])) % 65536
    assert [int(x) for x in (d / "io" / "toy" / "result.txt").read_text().split()] == [want]
