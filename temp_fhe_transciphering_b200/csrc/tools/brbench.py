"""Development tool: time the blind rotation alone and the whole circuit bootstrap (blind rotation + trace +
scheme switch) for one batch size, under whatever CBS_*_VARIANT environment is set.  usage: brbench.py [batch]"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import temp_fhe_transciphering_b200 as cbs

ks = cbs.KeySet.generate(1)
ctx = cbs.Context(ks, 0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ctx.set_stream(st.cuda_stream)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
small = torch.from_numpy(ks.encrypt_bits_small(np.random.default_rng(0).integers(0, 2, B, dtype=np.uint8), 5).view(np.int64)).cuda()
acc = torch.empty((B, 3072), dtype=torch.int64, device="cuda")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


br = timed(lambda: ctx.blind_rotate_dev(small.data_ptr(), acc.data_ptr(), B))
cb = timed(lambda: ctx.circuit_bootstrap_dev(small.data_ptr(), B))
print("B", B,
      "blind_rotate_ms %.3f" % br, "circuit_bootstrap_ms %.3f" % cb, "trace+ss_ms %.3f" % (cb - br))
