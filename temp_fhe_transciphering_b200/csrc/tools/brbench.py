import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle')
import numpy as np, torch, time
import temp_fhe_transciphering_b200 as cbs
ks = cbs.KeySet.generate(1)
ctx = cbs.Context(ks, 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); ctx.set_stream(st.cuda_stream)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
small = torch.from_numpy(ks.encrypt_bits_small(np.random.default_rng(0).integers(0,2,B,dtype=np.uint8), 5).view(np.int64)).cuda()
acc = torch.empty((B,3072), dtype=torch.int64, device='cuda')
for _ in range(2): ctx.blind_rotate_dev(small.data_ptr(), acc.data_ptr(), B)
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(5): ctx.blind_rotate_dev(small.data_ptr(), acc.data_ptr(), B)
e1.record(st); torch.cuda.synchronize()
print(os.environ.get('CBS_BR_VARIANT'), 'B', B, 'ms', e0.elapsed_time(e1)/5)
