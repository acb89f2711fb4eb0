import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from seqalib_b200 import capi, synth
n = 1_000_000
lib = capi.Lib()
prm = capi.make_params("sw", gap=-1, match=1, mismatch=-1)
ctx = capi.Ctx(lib); ctx.generate(prm, synth.SEED, 0, n, 0, 150, 150)
hb, o1, o2, l1, l2 = ctx.download_inputs(n * 300); ctx.close()
def pinned(shape, dt):
    t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dt).reshape(shape)
pb = pinned(n * 300, np.uint8); pb[:] = hb
po1 = pinned(n, np.uint64); po1[:] = o1
po2 = pinned(n, np.uint64); po2[:] = o2
pl1 = pinned(n, np.uint32); pl1[:] = l1
pl2 = pinned(n, np.uint32); pl2[:] = l2
res = capi.Results(n, n * 300, pinned=pinned)
os.environ["SEQA_WORKERS"] = sys.argv[1]
for k in range(4):
    if k == 3: os.environ["SEQA_DEBUG_TIMING"] = "1"
    t = time.perf_counter(); lib.align_batch(prm, pb, po1, po2, pl1, pl2, res); dt = time.perf_counter() - t
    print("align_batch call %d: %.2f ms" % (k, dt * 1e3))
