"""e2e knob sweep: seqa_cuda_align_batch wall time vs SEQA_RING / SEQA_WAVE_MCELLS for both input wire formats
(experiment harness; results summarised in profiles/)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqalib_b200 import capi, synth
n = 1_000_000
lib = capi.Lib()
def pinned(shape, dt):
    t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dt).reshape(shape)
pb = pinned(n * 300, np.uint8)
_, o1, o2, l1, l2 = synth.batch_uniform(synth.SEED, 0, n, 150, 150, out=pb)
po1 = pinned(n, np.uint64); po1[:] = o1
po2 = pinned(n, np.uint64); po2[:] = o2
pl1 = pinned(n, np.uint32); pl1[:] = l1
pl2 = pinned(n, np.uint32); pl2[:] = l2
pk = pinned(n * 76, np.uint8)
_, k1, k2 = capi.pack_bases_2bit(pb, po1, po2, pl1, pl2, out=pk)
pk1 = pinned(n, np.uint64); pk1[:] = k1
pk2 = pinned(n, np.uint64); pk2[:] = k2
res = capi.Results(n, n * 300, pinned=pinned)
rings = [int(x) for x in os.environ.get("SWEEP_RINGS", "4").split(",")]
waves = [int(x) for x in os.environ.get("SWEEP_WAVES", "2600,5100,7700,10200").split(",")]
for name, flags, ins in (("2bit-in", capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT, (pk, pk1, pk2, pl1, pl2)),
                         ("8bit-in", capi.FLAG_OPS_2BIT, (pb, po1, po2, pl1, pl2))):
    prm = capi.make_params("sw", gap=-1, match=1, mismatch=-1, flags=flags)
    for ring in rings:
        for mcells in waves:
            os.environ["SEQA_RING"] = str(ring); os.environ["SEQA_WAVE_MCELLS"] = str(mcells)
            ts = []
            for k in range(8):
                t = time.perf_counter(); lib.align_batch(prm, *ins, res); ts.append(time.perf_counter() - t)
            print("%s ring %d wave %5d Mcells: best %.2f ms  median %.2f ms" % (name, ring, mcells, min(ts[2:]) * 1e3, float(np.median(ts[2:])) * 1e3), flush=True)
            lib.L.seqa_cuda_trim()
