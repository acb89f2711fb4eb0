"""Kernel / memcpy timeline of ONE end-to-end seqa_cuda_align_batch call (2-bit wire formats), through CUPTI (torch.profiler):
start, duration and stream of every kernel and copy, in ms after the first activity.  usage: e2e_trace.py [pairs]"""
import os, sys
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqalib_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lib = capi.Lib()
def pinned(shape, dt):
    t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dt).reshape(shape)
pb = pinned(n * 300, np.uint8)
_, o1, o2, l1, l2 = synth.batch_uniform(synth.SEED, 0, n, 150, 150, out=pb)
pl1 = pinned(n, np.uint32); pl1[:] = l1
pl2 = pinned(n, np.uint32); pl2[:] = l2
pk = pinned(n * 76, np.uint8)
_, k1, k2 = capi.pack_bases_2bit(pb, o1, o2, pl1, pl2, out=pk)
pk1 = pinned(n, np.uint64); pk1[:] = k1
pk2 = pinned(n, np.uint64); pk2[:] = k2
res = capi.Results(n, n * 300, pinned=pinned)
prm = capi.make_params("sw", gap=-1, match=1, mismatch=-1, flags=capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT)
for _ in range(3):
    lib.align_batch(prm, pk, pk1, pk2, pl1, pl2, res)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    lib.align_batch(prm, pk, pk1, pk2, pl1, pl2, res)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print("%-44s %9s %9s  stream-ish" % ("activity", "start ms", "dur ms"))
for e in ev:
    d = (e.time_range.end - e.time_range.start) / 1e3
    if d < 0.004 and "Memcpy" not in e.name and "walk" not in e.name:
        continue
    print("%-44s %9.3f %9.3f" % (e.name[:44], (e.time_range.start - t0) / 1e3, d))
