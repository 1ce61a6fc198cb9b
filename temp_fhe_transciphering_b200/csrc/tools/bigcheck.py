"""Development tool: transcipher many blocks twice with seeded keys and report (a) wrong blocks against the cleartext,
(b) whether the two runs are bit-identical (a difference = a race, not noise), (c) the output-noise tail.
usage: bigcheck.py [blocks] [mode ecb|ctr] [seed]"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import aes_clear
import ref_io
import temp_fhe_transciphering_b200 as cbs

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "ctr"
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ks = cbs.KeySet.generate(seed)
ctx = cbs.Context(ks, 0)
key = aes_clear.harness_aes_key(None)
pt = bytes(np.random.default_rng(seed).integers(0, 256, 16 * nb, dtype=np.uint8))
outs = []
for run in range(2):
    if mode == "ctr":
        iv = aes_clear.harness_iv(None)
        ct = aes_clear.ctr_crypt(key, iv, pt)
        kf = ks.gen_forward_transciphering_keys(key, seed + 1)
        got = ctx.aes_ctr_to_lwe_transciphering(ct, iv, *kf)
    else:
        ct = aes_clear.ecb_encrypt(key, pt)
        tk = ks.gen_transciphering_keys(key, seed + 1)
        got = ctx.aes_to_lwe_transciphering(ct, *tk)
    outs.append(np.array(got).reshape(-1, 2049).copy())
same = bool((outs[0] == outs[1]).all())
ph = ref_io.lwe_phase(outs[0], ks.glwe_sk)
bits = ref_io.decode_bit(ph)
dec = np.packbits(bits).tobytes()
wrong_blocks = [b for b in range(nb) if dec[16 * b:16 * b + 16] != pt[16 * b:16 * b + 16]]
want_bits = np.unpackbits(np.frombuffer(pt, dtype=np.uint8))
err = ref_io.bit_error(ph, want_bits)  # distance from the CORRECT bit
a = np.abs(err)
print({"blocks": nb, "mode": mode, "seed": seed, "runs_bit_identical": same, "wrong_blocks": wrong_blocks,
       "wrong_bits": int((bits != want_bits).sum()), "noise_log2_std": float(np.log2(np.sqrt(np.mean(err[a < 2.0**62] ** 2)))),
       "noise_log2_max": float(np.log2(a.max())), "tail_counts_over_2^60/60.5/61/61.5": [int((a > 2.0**e).sum()) for e in (60, 60.5, 61, 61.5)]})
if not same:
    d = np.nonzero((outs[0] != outs[1]).any(axis=1))[0]
    print("rows differing between runs:", len(d), d[:20], "blocks", sorted(set((d // 128).tolist()))[:20])
