#!/usr/bin/env python
"""Pins the CPU oracle against the REFERENCE ITSELF (run here, where /root/reference exists).

The reference has no unit tests, golden vectors or KATs for this path; its FHE
key generation is unseeded, so only decrypted outputs are comparable.  This script creates the pins:

 A. reference keys -> oracle.  The reference's own prebuilt client binaries generate keys and the
    transciphering key; the reference's stage-7/stage-8 binaries and the oracle both run on those
    files; both results are decrypted with the reference's secret key.
 B. seeded keys -> reference.  Keys and transciphering key come from OUR seeded client helpers
    (written in the reference's bincode formats); the unmodified reference stage-7 binary runs on them.
    Because the inputs are a pure function of the seeds below, the recorded reference output
    statistics are a reproducible golden vector: tests/test_oracle_cpu.py regenerates the same inputs
    on any machine and checks the oracle (and, on the GPU box, the CUDA path) against it.

Writes tests/golden/reference_pin.json.  Takes ~3 minutes (two ~45 s single-threaded reference runs).
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import aes_clear  # noqa: E402
import oracle  # noqa: E402
import ref_io  # noqa: E402

REF = os.path.join(HERE, "_ref")
PIN_SEED_KEYS = 1
PIN_SEED_TRANS = 7
PIN_PLAINTEXT = bytes(range(16))


def run(binary, cwd):
    t0 = time.time()
    subprocess.run([os.path.join(REF, binary), "0"], cwd=cwd, check=True, stdout=subprocess.DEVNULL)
    return time.time() - t0


def stats(lwe, sk, expect_bytes):
    bits, std, mx = ref_io.noise_stats(lwe, sk)
    ph = ref_io.lwe_phase(lwe, sk)
    err = ref_io.bit_error(ph, bits)
    return {"bytes_hex": np.packbits(bits).tobytes().hex(), "correct": np.packbits(bits).tobytes() == expect_bytes,
            "noise_log2_std": std, "noise_log2_max": mx, "phase_error_int64": [int(e) for e in err]}


def main():
    import temp_fhe_transciphering_b200 as cbs
    oracle.build()
    pin = {"generated_by": "oracle/pin_against_reference.py", "reference_binaries": "submission/target/release (rustc 1.90, tfhe 0.5.4)"}
    aes_key = aes_clear.harness_aes_key(None)

    # ---------------- A: reference keys -> reference server + oracle ----------------
    with tempfile.TemporaryDirectory() as d:
        vals = [20962, 11749, 64797, 2177, 19876, 44457, 4094, 20862]
        pt = aes_clear.pack_u16_be(vals)
        os.makedirs(f"{d}/datasets/toy")
        open(f"{d}/datasets/toy/aes_key.hex", "w").write(aes_key.hex())
        open(f"{d}/datasets/toy/db.hex", "w").write(aes_clear.ecb_encrypt(aes_key, pt).hex())
        run("client_key_generation", d)
        run("client_encode_encrypt", d)
        t7 = run("server_encrypted_aes_decryption", d)
        t8 = run("server_encrypted_compute", d)
        io = f"{d}/io/toy"
        sk = ref_io.read_lwe_sk(f"{io}/secret_keys/lwe_sk.bin")
        inp = ref_io.load_server_inputs(io)
        ref7 = ref_io.read_lwe_list(f"{io}/ciphertext_aes_download/result.bin")
        ref8 = ref_io.read_lwe_list(f"{io}/ciphertexts_download/result.bin")
        K = oracle.Keys(inp["bsk"], inp["ksk"], inp["auto_std"], inp["ss"])
        t0 = time.time()
        orc7 = oracle.aes128_transcipher(K, ref_io.read_db_hex(f"{d}/datasets/toy/db.hex"), *inp["trans_key"])[0]
        to7 = time.time() - t0
        t0 = time.time()
        orc8 = oracle.max_u16(K, ref7)
        to8 = time.time() - t0
        mx = aes_clear.pack_u16_be([max(vals)])
        a = {"values": vals, "reference_stage7": stats(ref7, sk, pt), "oracle_stage7": stats(orc7, sk, pt),
             "reference_stage8": stats(ref8, sk, mx), "oracle_stage8": stats(orc8, sk, mx),
             "reference_stage7_seconds_1core": t7, "reference_stage8_seconds_1core": t8,
             "oracle_stage7_seconds": to7, "oracle_stage8_seconds": to8, "oracle_threads": oracle.num_threads(),
             "auto_keys_roundtrip_err": inp["auto_roundtrip_err"]}
        for k in ("reference_stage7", "oracle_stage7", "reference_stage8", "oracle_stage8"):
            a[k].pop("phase_error_int64")  # unseeded keys: not reproducible, keep the statistics only
        pin["A_reference_keys"] = a

    # ---------------- B: seeded keys -> reference server ----------------
    with tempfile.TemporaryDirectory() as d:
        ks = cbs.KeySet.generate(PIN_SEED_KEYS)
        tk = ks.gen_transciphering_keys(aes_key, PIN_SEED_TRANS)
        ct = aes_clear.ecb_encrypt(aes_key, PIN_PLAINTEXT)
        ks.save_dir(f"{d}/io/toy", with_secret=True)
        cbs.save_trans_key(f"{d}/io/toy/ciphertexts_upload/trans_key.bin", *tk)
        os.makedirs(f"{d}/datasets/toy")
        open(f"{d}/datasets/toy/db.hex", "w").write(ct.hex())
        t7 = run("server_encrypted_aes_decryption", d)
        ref7 = ref_io.read_lwe_list(f"{d}/io/toy/ciphertext_aes_download/result.bin")
        K = oracle.Keys(ks.bsk, ks.ksk, ks.auto_std, ks.ss)
        orc7 = oracle.aes128_transcipher(K, ct, *tk)[0]
        pin["B_seeded_keys"] = {
            "seed_keys": PIN_SEED_KEYS, "seed_trans_key": PIN_SEED_TRANS, "aes_key_hex": aes_key.hex(),
            "plaintext_hex": PIN_PLAINTEXT.hex(), "ciphertext_hex": ct.hex(),
            "reference_stage7": stats(ref7, ks.glwe_sk, PIN_PLAINTEXT),
            "oracle_stage7": stats(orc7, ks.glwe_sk, PIN_PLAINTEXT),
            "reference_stage7_seconds_1core": t7,
            "keyset_checksum": int(np.bitwise_xor.reduce(ks.bsk.reshape(-1))) ^ int(np.bitwise_xor.reduce(ks.auto_std.reshape(-1))),
        }
    out = os.path.join(ROOT, "tests", "golden", "reference_pin.json")
    json.dump(pin, open(out, "w"), indent=1)
    a, b = pin["A_reference_keys"], pin["B_seeded_keys"]
    print("A stage7 ref/oracle std:", a["reference_stage7"]["noise_log2_std"], a["oracle_stage7"]["noise_log2_std"],
          "correct:", a["reference_stage7"]["correct"], a["oracle_stage7"]["correct"])
    print("A stage8 ref/oracle std:", a["reference_stage8"]["noise_log2_std"], a["oracle_stage8"]["noise_log2_std"],
          "correct:", a["reference_stage8"]["correct"], a["oracle_stage8"]["correct"])
    print("B stage7 ref/oracle std:", b["reference_stage7"]["noise_log2_std"], b["oracle_stage7"]["noise_log2_std"],
          "correct:", b["reference_stage7"]["correct"], b["oracle_stage7"]["correct"])


if __name__ == "__main__":
    main()
