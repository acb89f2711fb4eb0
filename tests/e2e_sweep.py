"""e2e knob sweep: seqa_cuda_align_batch wall time vs SEQA_RING / SEQA_WAVE_MCELLS (experiment harness)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqalib_b200 import capi, synth
n = 1_000_000
lib = capi.Lib()
prm = capi.make_params("sw", gap=-1, match=1, mismatch=-1)
prm2 = capi.make_params("sw", gap=-1, match=1, mismatch=-1, flags=capi.FLAG_OPS_2BIT)
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
prm = prm2
for ring in (3, 4, 5):
    for mcells in (2500, 5000):
        os.environ["SEQA_RING"] = str(ring); os.environ["SEQA_WAVE_MCELLS"] = str(mcells)
        ts = []
        for k in range(6):
            t = time.perf_counter(); lib.align_batch(prm, pb, po1, po2, pl1, pl2, res); ts.append(time.perf_counter() - t)
        print("ring %d wave %5d Mcells: best %.2f ms  median %.2f ms" % (ring, mcells, min(ts[2:]) * 1e3, float(np.median(ts[2:])) * 1e3), flush=True)
        lib.L.seqa_cuda_trim()
os.environ["SEQA_RING"] = "3"; os.environ["SEQA_WAVE_MCELLS"] = "3000"
lib.align_batch(prm, pb, po1, po2, pl1, pl2, res)
os.environ["SEQA_DEBUG_TIMING"] = "1"
lib.align_batch(prm, pb, po1, po2, pl1, pl2, res)
