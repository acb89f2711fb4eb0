"""Where does an end-to-end seqa_cuda_align_batch call spend its time?  (PCIe copy rates + per-wave device timeline,
SEQA_DEBUG_TIMING, for the 8-bit and the 2-bit input wire formats)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqalib_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x = torch.empty(324_000_000, dtype=torch.uint8, pin_memory=True); d = torch.empty_like(x, device="cuda")
for _ in range(2): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t = time.perf_counter(); d.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("H2D pinned 324 MB: %.2f ms = %.1f GB/s" % (dt * 1e3, 0.324 / dt))
t = time.perf_counter(); x.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("D2H pinned 324 MB: %.2f ms = %.1f GB/s" % (dt * 1e3, 0.324 / dt))
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
os.environ.pop("SEQA_DEBUG_TIMING", None)
for name, flags, ins in (("8-bit symbols in, 2-bit ops out", capi.FLAG_OPS_2BIT, (pb, po1, po2, pl1, pl2)),
                         ("2-bit symbols in, 2-bit ops out", capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT, (pk, pk1, pk2, pl1, pl2))):
    prm = capi.make_params("sw", gap=-1, match=1, mismatch=-1, flags=flags)
    print("==== %s" % name, flush=True)
    for k in range(5):
        if k == 4:
            os.environ["SEQA_DEBUG_TIMING"] = "1"
        t = time.perf_counter(); lib.align_batch(prm, *ins, res); dt = time.perf_counter() - t
        print("align_batch call %d: %.2f ms" % (k, dt * 1e3), flush=True)
    os.environ.pop("SEQA_DEBUG_TIMING", None)
