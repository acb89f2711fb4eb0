"""TEST INFRASTRUCTURE.  Golden fingerprints of BASELINE configs[3] at FULL size: the 64 pairs of 100,000 x 100,000 bp
random DNA (shared generator, seed 20240607, pairs 0..63) aligned by the C oracle (oracle/seqa_oracle.c, itself pinned
against the compiled reference) with HirschbergSA (-1,2,-1) and MyersMillerSA (-3,-1,1,-1).  Per pair: score, ops_len
and CRC-32 of the op string (one byte per op, forward order).  bench.py's config-4 entry and the -m gpu tests compare
every pair the GPU aligned with these fingerprints: a full-size, all-pairs parity check that needs no CPU time on
the GPU box (the oracle takes ~38 s / ~165 s per pair and core).

    python oracle/make_golden_config4.py [threads]      # rewrites tests/golden/config4_100kbp.json  (~30 min on 8 cores)
"""
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as orc  # noqa: E402
from seqalib_b200 import synth  # noqa: E402

N_PAIRS, L = 64, 100_000
CASES = [("hirschberg", orc.Scoring.linear(-1, 2, -1)), ("myersmiller", orc.Scoring.affine(-3, -1, 1, -1))]


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else len(os.sched_getaffinity(0))
    orc.build()
    bases, off1, off2, l1, l2 = synth.batch(synth.SEED, 0, N_PAIRS, 0, L, L)
    out = {"seed": synth.SEED, "first_pair": 0, "pairs": N_PAIRS, "len": L, "generator": "seqalib_b200/synth.py: batch(seed, 0, 64, 0, L, L)",
           "oracle": "oracle/seqa_oracle.c via oracle_align_batch", "algos": {}}
    for algo, sc in CASES:
        t0 = time.time()
        r = orc.oracle_align_batch(algo, sc, bases, off1, off2, l1, l2, threads=threads)
        rows = []
        for p in range(N_PAIRS):
            ops = np.ascontiguousarray(r.pair_ops(p))
            rows.append({"score": int(r.score[p]), "ops_len": int(r.ops_len[p]), "crc32": zlib.crc32(ops.tobytes()) & 0xffffffff})
        out["algos"][algo] = {"scoring": list(sc.astuple()), "pairs": rows}
        sys.stderr.write("%s: %d pairs in %.0f s\n" % (algo, N_PAIRS, time.time() - t0))
    dst = os.path.join(ROOT, "tests", "golden", "config4_100kbp.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=0)
        f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
