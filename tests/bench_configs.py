"""Secondary measurements for the BASELINE configs that are not the bench.py headline (configs[2], [3], [4]) at
sizes that fit a few GPU-minutes: device-resident GCUPS (seqa_ctx_run, CUDA-event fill time + wall time per run)
next to the reference CPU path on a small sample.  Output: one JSON object per line; summarised in profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seqalib_b200 import capi, synth  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

W = {"nw": 7, "sw": 10, "ggotoh": 11, "lgotoh": 14, "hirschberg": 14, "myersmiller": 22}


def cpu_sample(algo, sc, n, len_mode, l1, l2, first=0):
    bases, off1, off2, a, b = synth.batch(synth.SEED, first, n, len_mode, l1, l2)
    threads = len(os.sched_getaffinity(0))
    if orc.have_ref():
        sec, _ = orc.ref_bench(algo, sc, bases, off1, off2, a, b, threads)
        kind = "reference"
    else:
        sec, _ = orc.oracle_bench(algo, sc, bases, off1, off2, a, b, threads)
        kind = "port"
    return float((a.astype(np.float64) * b).sum()) / sec / 1e9, threads, kind


def run(lib, name, algo, sc, n, len_mode, l1, l2, cpu_n, reps=3, related=False):
    prm = capi.make_params(algo, gap=sc.gap, gap_open=sc.gap_open, gap_extend=sc.gap_extend, match=sc.match,
                           mismatch=sc.mismatch if sc.allow else 0, allow=sc.allow)
    ctx = capi.Ctx(lib)
    if related:  # config 4's realism variant: seq2 derived from seq1 (synth.related_sequence), uploaded once
        ctx.upload(prm, *synth.related_batch(synth.SEED, 0, n, l1))
    else:
        ctx.generate(prm, synth.SEED, 0, n, len_mode, l1, l2)
    ctx.run()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.run()
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    fill_ms, nl = ctx.last_fill_ms()
    cells = ctx.cells()
    f_clk = 1.965e9
    out = {"config": name, "algo": algo, "pairs": n, "cells": cells, "kernel": ctx.last_kernel(),
           "gcups_step": cells / dt / 1e9, "ms_step": dt * 1e3, "gcups_fill": cells / (fill_ms * 1e-3) / 1e9,
           "fill_ms": fill_ms, "fill_launches": nl,
           "roofline_frac_fill": cells * W[algo] / (fill_ms * 1e-3) / (148 * 128 * f_clk)}
    if cpu_n:
        g, thr, kind = cpu_sample(algo, sc, cpu_n, len_mode, l1, l2)
        out["cpu_gcups"] = g
        out["cpu_cores"] = thr
        out["cpu_kind"] = kind
        out["cpu_sample_pairs"] = cpu_n
    ctx.close()
    print(json.dumps(out), flush=True)


def main():
    lib = capi.Lib()
    S = orc.Scoring
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    only = sys.argv[2] if len(sys.argv) > 2 else ""   # substring filter on the config name
    global run
    run_all = run

    def run(lib, name, *a, **k):
        if only in name:
            run_all(lib, name, *a, **k)
    run(lib, "config2 NW 150bp", "nw", S.linear(-1, 2, -1), int(1_000_000 * scale), 0, 150, 150, 20000)
    run(lib, "config3 GlobalGotoh 250bp", "ggotoh", S.affine(-3, -1, 1, -1), int(200_000 * scale), 0, 250, 250, 8000)
    run(lib, "config3 LocalGotoh 250bp", "lgotoh", S.affine(-3, -1, 1, -1), int(200_000 * scale), 0, 250, 250, 8000)
    run(lib, "config5 mixed 50-1000bp NW", "nw", S.linear(-1, 2, -1), int(500_000 * scale), 1, 0, 0, 2000)
    run(lib, "config5 mixed 50-1000bp SW", "sw", S.linear(-1, 1, -1), int(500_000 * scale), 1, 0, 0, 2000)
    run(lib, "config4 Hirschberg 20kbp x16", "hirschberg", S.linear(-1, 2, -1), 16, 0, 20000, 20000, 0, reps=1)
    run(lib, "config4 MyersMiller 20kbp x16", "myersmiller", S.affine(-3, -1, 1, -1), 16, 0, 20000, 20000, 0, reps=1)
    run(lib, "config4 Hirschberg 100kbp x8", "hirschberg", S.linear(-1, 2, -1), 8, 0, 100000, 100000, 0, reps=1)
    run(lib, "config4 Hirschberg 100kbp x64", "hirschberg", S.linear(-1, 2, -1), 64, 0, 100000, 100000, 0, reps=1)
    run(lib, "config4 MyersMiller 100kbp x64", "myersmiller", S.affine(-3, -1, 1, -1), 64, 0, 100000, 100000, 0, reps=1)
    run(lib, "config4 related Hirschberg 100kbp x64", "hirschberg", S.linear(-1, 2, -1), 64, 0, 100000, 100000, 0, reps=1, related=True)
    run(lib, "config4 related MyersMiller 100kbp x64", "myersmiller", S.affine(-3, -1, 1, -1), 64, 0, 100000, 100000, 0, reps=1, related=True)


if __name__ == "__main__":
    main()
