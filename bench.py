#!/usr/bin/env python
"""bench.py -- headline benchmark of the SeqALib DP hot path on B200 (BASELINE.json configs[1]):

    Smith-Waterman linear gap, ScoringSystem(-1,1,-1), 1,000,000 random DNA pairs of 150 bp PER GPU,
    score + traceback, GCUPS = sum(len1*len2) / seconds / 1e9.

A "step" = one pass of the hot path over the whole batch.  `value` times it with the inputs resident in HBM
(seqa_ctx_run: prep + fill + walk + scan + gather); `e2e` times the reference-facing C-ABI call
seqa_cuda_align_batch with pinned HOST buffers in and out.  One process per GPU (torchrun for N > 1), pairs
sharded statically, no data-path collective: weak scaling.  `--impl reference` times the reference's own CPU
implementation (oracle/_ref: the unmodified reference headers compiled by path; else the C port) on the host.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 1_000_000
LEN = 150
SEED = 20240607
SCORING = dict(gap=-1, match=1, mismatch=-1)  # SmithWatermanSA default (reference include/SASmithWaterman.h:352)
W_OPS_PER_CELL = 10   # algorithmic INT32 ops per SW cell
ALG_TRACE_BITS = 2    # algorithmic traceback bits per cell (linear gap)
METRIC = "GCUPS (score+traceback) batched 150bp SW/NW at 1/2/4/8 B200 vs host CPU"


_JSON_OUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner, torch warnings)
    is sent to stderr instead, and emit() writes to the saved descriptor."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def kernel_traffic(kernel, pairs):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
    (profiles/r02_traffic.json, else r01: dram__bytes_read.sum + dram__bytes_write.sum), scaled by the pair count.
    The contract allows the capture to be a separate run (a number printed under ncu is never a bench value)."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)[kernel]
            return float(t["dram_bytes_per_launch"]) * pairs / float(t["pairs"])
        except Exception:
            continue
    return None


def sass_alu_per_two_cells(kernel):
    """ALU-pipe instructions per two cells in the hot loop of `kernel`, counted from the SASS of the committed build by
    tests/sass_count.py (profiles/sass_counts.json).  Falls back to the hand count of seqa_packed.cuh (5.25)."""
    try:
        with open(os.path.join(ROOT, "profiles", "sass_counts.json")) as f:
            t = json.load(f)[kernel]
        return float(t["alu_per_2_cells"]), "profiles/sass_counts.json (%s)" % t.get("function", kernel)
    except Exception:
        return 5.25, "hand count (seqa_packed.cuh header comment)"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=8.0):
        """block until nvidia-smi has delivered its first sample (its start-up can outlast a 50 ms timed region)"""
        t_end = time.time() + timeout
        while self.proc and not self.rows and time.time() < t_end:
            time.sleep(0.01)

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        out = self.window(t0, t1)
        self.close()
        return out

    def close(self):
        if self.proc:
            self.proc.terminate()
            self.proc = None

    def window(self, t0, t1):
        """clocks / throttle reasons of the samples taken between two mark()s (the sampler keeps running)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        sm, mx, reasons = [], None, set()
        for (t, line) in list(self.rows):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk = float(f[1])
                mx = float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: take the nearest samples
            sm = [float(r[1].split(",")[1]) for r in self.rows[-3:] if len(r[1].split(",")) > 2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


_CPU_BATCH = {}


def cpu_reference_gcups(n_pairs, threads):
    """The reference's own CPU path (SmithWatermanSA::getAlignment with equal<char>, one aligner object per
    std::thread) on `n_pairs` pairs of the bench workload.  -> (gcups, kind, seconds)"""
    from oracle import pyoracle as orc
    from seqalib_b200 import synth
    if not os.path.exists(orc.ORACLE_SO):
        orc.build()
    if _CPU_BATCH.get("n") != n_pairs:  # generated once (vectorised), reused by every step
        _CPU_BATCH.clear()
        _CPU_BATCH.update(n=n_pairs, arrays=synth.batch_uniform(SEED, 0, n_pairs, LEN, LEN))
    bases, off1, off2, l1, l2 = _CPU_BATCH["arrays"]
    sc = orc.Scoring.linear(SCORING["gap"], SCORING["match"], SCORING["mismatch"])
    cells = float((l1.astype(np.float64) * l2).sum())
    if orc.have_ref():
        sec, _ = orc.ref_bench("sw", sc, bases, off1, off2, l1, l2, threads)
        kind = "reference"
    else:
        sec, _ = orc.oracle_bench("sw", sc, bases, off1, off2, l1, l2, threads)
        kind = "port"
    return cells / sec / 1e9, kind, sec


def header_api_bench(pairs, device_index=0):
    """e2e.api_packed / e2e.api_list: tests/cpp/bench_header.cpp (the reference's own class templates from
    include/SequenceAlignment.h, std::string pairs in pageable memory) compiled against the in-tree library and run as a
    separate process on this GPU.  -> {"packed": {...}, "list": {...}, "note": ...} or {"note": why not}"""
    import tempfile
    exe = os.path.join(tempfile.mkdtemp(prefix="seqa_bench_"), "bench_header")
    libdir = os.path.join(ROOT, "seqalib_b200")
    try:
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", "bench_header.cpp"), "-o", exe, "-L", libdir, "-lseqa_cuda",
                               "-Wl,-rpath," + libdir], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=300)
        # the header uses every VISIBLE device: show it this rank's GPU only, so that the figure belongs to the N = 1 line
        env = dict(os.environ)
        vis = [v for v in env.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        env["CUDA_VISIBLE_DEVICES"] = vis[device_index] if device_index < len(vis) else str(device_index)
        out = subprocess.run([exe, str(pairs), "3", str(min(pairs, 200000))], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
        d = json.loads(line)
        d["note"] = ("getAlignmentsPacked on %d and getAlignments on %d std::string pairs of %d bp, best of 3 / 2 calls after a warm-up call, "
                     "%d host threads; packing, PCIe and (list) std::list construction inside the timed call"
                     % (d["packed"]["pairs"], d["list"]["pairs"], LEN, d.get("host_threads", 0)))
        return d
    except Exception as e:  # no compiler on the box, ...
        return {"note": "template-API legs unavailable: %r" % (e,)}


def bind_to_gpu_numa_node(index):
    """Multi-GPU runs: keep this rank's threads and its pinned host buffers on the NUMA node its GPU hangs off
    (torchrun does not place ranks; a remote node halves the PCIe copy rate of the end-to-end leg)."""
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                           stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip()
        dom, bus, rest = q.lower().split(":")
        node = int(open("/sys/bus/pci/devices/%s:%s:%s/numa_node" % (dom[-4:], bus, rest)).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    # the CPU arm on the FULL 1 M-pair set when the whole --steps/--warmup run still ends within a few
    # minutes on this host (~0.18 GCUPS per thread), else a bounded sample of the same generator stream
    per_step_s = args.pairs * LEN * LEN / (threads * 0.18e9)
    if per_step_s * (args.steps + args.warmup) <= 330:
        n = args.pairs
    else:
        n = max(8000, int(args.pairs * 330 / (per_step_s * (args.steps + args.warmup))))
    vals, kind = [], "port"
    for k in range(args.warmup + args.steps):
        g, kind, sec = cpu_reference_gcups(n, threads)
        if k >= args.warmup:
            vals.append((g, sec))
    value = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    sample = ("the full %d pairs of %d bp per step (same generator/seed as the GPU arm)" if n == args.pairs else
              "%d pairs of %d bp per step (same generator/seed as the GPU arm)") % (n, LEN)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            # the same workload name as the GPU arm's line; the CPU arm times a bounded sample of it per step
            "config": {"workload": "SmithWatermanSA linear gap (-1,1,-1), %d random DNA pairs of %d bp per GPU, score+traceback" % (args.pairs, LEN),
                       "pairs_per_gpu": args.pairs, "len": LEN, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# The other BASELINE configs (configs[2..4]) in the same JSON line: "configs": [...].  Strong scaling: the config's
# whole workload is split over the N ranks (no collective on the data path), every rank times its own share with CUDA
# events on its stream, the entry reports total cells / max-over-ranks time.
W_OPS = {"nw": 7, "sw": 10, "ggotoh": 11, "lgotoh": 14, "hirschberg": 14, "myersmiller": 22}


def _spot_check(ctx, algo, sc, seed, first_pair, seg_n, len_mode, l1, l2, want, rng):
    """`want` pairs of the resident segment against the oracle (the checker, outside every timed region): the pair is
    regenerated on the CPU from the shared counter-based generator, its results are read back with
    seqa_ctx_download_range.  -> pairs checked (raises on any difference)."""
    from oracle import pyoracle as orc
    from seqalib_b200 import synth
    if want <= 0 or seg_n == 0:
        return 0
    picks = sorted(set([0, seg_n - 1] + [int(x) for x in rng.integers(0, seg_n, max(want - 2, 0))]))[:want]
    for k in picks:
        pid = first_pair + k
        if len_mode:
            a_len, b_len = synth.lengths(seed, pid, 1)
            a_len, b_len = int(a_len[0]), int(b_len[0])
        else:
            a_len, b_len = l1, l2
        a = bytes(synth.sequence(seed, pid, 0, a_len)).decode()
        b = bytes(synth.sequence(seed, pid, 1, b_len)).decode()
        r = ctx.download_range(k, 1, a_len + b_len + 8)
        o = orc.oracle_align(algo, sc, a, b)
        got = (int(r.score[0]), int(r.start_i[0]), int(r.start_j[0]), int(r.end_i[0]), int(r.end_j[0]))
        exp = (o["score"], o["start_i"], o["start_j"], o["end_i"], o["end_j"])
        if got != exp or not np.array_equal(r.pair_ops(0), o["ops"]):
            raise SystemExit("bench.py spot check FAILED: %s pair %d: got %r, oracle %r" % (algo, pid, got, exp))
    return len(picks)


def _resident_pass(lib, capi, torch, local, stream, algo, sc, seed, segments, len_mode, l1, l2, spot, fingerprints=None):
    """One pass of `algo` over this rank's share, given as a list of (first_pair, n_pairs) segments that are generated
    on the device one after another (a share larger than HBM is streamed through in segments; the host-side planning
    of a segment is outside the timed region, like the resident `value` of the headline).  The first segment is run
    once untimed (allocations, instruction cache).  -> dict(ms, fill_ms, cells, launches, kernel, checked)"""
    prm = capi.make_params(algo, gap=sc.gap, gap_open=sc.gap_open, gap_extend=sc.gap_extend, match=sc.match,
                           mismatch=sc.mismatch if sc.allow else 0, allow=sc.allow, device_first=local, device_count=1)
    ctx = capi.Ctx(lib, local, stream.cuda_stream)
    out = {"ms": 0.0, "fill_ms": 0.0, "cells": 0, "launches": 0, "kernel": "none", "checked": 0, "pairs": 0}
    rng = np.random.default_rng(12345 + local)
    warmed = False
    for (first, n) in segments:
        if n == 0:
            continue
        ctx.generate(prm, seed, first, n, len_mode, l1, l2)
        if not warmed:
            ctx.run()
            ctx.sync()
            warmed = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count()
        e0.record(stream)
        ctx.run()
        e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        out["ms"] += e0.elapsed_time(e1)
        fm, _ = ctx.last_fill_ms()
        out["fill_ms"] += fm
        out["cells"] += ctx.cells()
        out["launches"] += ctx.launch_count() - l0
        out["kernel"] = ctx.last_kernel()
        out["pairs"] += n
        if fingerprints is not None:  # every pair of the segment against committed oracle fingerprints (config 4)
            import zlib
            res = ctx.download(ops_capacity=n * (l1 + l2))
            for k in range(n):
                g = fingerprints[first + k]
                ops = np.ascontiguousarray(res.pair_ops(k))
                got = (int(res.score[k]), int(res.ops_len[k]), zlib.crc32(ops.tobytes()) & 0xffffffff)
                if got != (g["score"], g["ops_len"], g["crc32"]):
                    raise SystemExit("bench.py fingerprint check FAILED: %s pair %d: got %r, golden %r" % (algo, first + k, got, g))
            out["checked"] += n
        elif spot:
            take = min(spot - out["checked"], max(2, spot // max(len(segments), 1) + 1))
            out["checked"] += _spot_check(ctx, algo, sc, seed, first, n, len_mode, l1, l2, take, rng)
    ctx.close()
    return out


def _config5_cuts(seed, n_total, world, block=16384, chunk=1 << 20):
    """Static split of the mixed-length stream balanced by sum(len1*len2): cumulative cells per block of
    16,384 pairs over the WHOLE stream (every rank computes the same numbers from the counter-based generator), cut at
    the block boundaries nearest to k/N of the total.  -> (pair cut points [world+1], cells per rank)"""
    from seqalib_b200 import synth
    from concurrent.futures import ThreadPoolExecutor
    nblocks = (n_total + block - 1) // block
    cells = np.zeros(nblocks, dtype=np.float64)

    def part(lo):
        hi = min(n_total, lo + chunk)
        a, b = synth.lengths(seed, lo, hi - lo)
        idx = (np.arange(lo, hi) // block).astype(np.int64)
        return np.bincount(idx, weights=a.astype(np.float64) * b, minlength=nblocks)
    with ThreadPoolExecutor(max_workers=max(1, min(4, host_threads() // max(world, 1)))) as ex:  # numpy releases the GIL
        for c in ex.map(part, range(0, n_total, chunk)):
            cells += c
    cum = np.concatenate([[0.0], np.cumsum(cells)])
    ks = [0]
    for r in range(1, world):
        ks.append(max(ks[-1], int(np.argmin(np.abs(cum - cum[-1] * r / world)))))
    ks.append(nblocks)
    cuts = [min(k * block, n_total) for k in ks]
    per_rank = [float(cum[ks[r + 1]] - cum[ks[r]]) for r in range(world)]
    return cuts, per_rank


def run_secondary_configs(args, lib, capi, torch, dist, local, rank, world, stream, sampler, f_max_mhz):
    from oracle import pyoracle as orc  # the checker for the spot checks; never inside a timed region
    from seqalib_b200 import shard
    S = orc.Scoring

    def gather(x):
        if world == 1:
            return [float(x)]
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        outl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outl, t)
        return [float(v.item()) for v in outl]

    def barrier():
        if world > 1:
            dist.barrier()

    entries = []

    def entry(name, algo, sc, segments, len_mode, l1, l2, spot, note, fingerprints=None, total_pairs=None):
        barrier()
        torch.cuda.synchronize()
        t0 = sampler.mark()
        r = _resident_pass(lib, capi, torch, local, stream, algo, sc, SEED, segments, len_mode, l1, l2, spot, fingerprints)
        t1 = sampler.mark()
        clocks = sampler.window(t0, t1)
        barrier()
        ms_all = gather(r["ms"])
        fill_all = gather(r["fill_ms"])
        cells_all = gather(r["cells"])
        checked_all = gather(r["checked"])
        launches = sum(gather(r["launches"]))
        if rank != 0:
            return
        ms = max(ms_all)
        tot_cells = sum(cells_all)
        f_clk = (clocks.get("sm_mhz") or f_max_mhz) * 1e6
        p_int = 148 * 128 * f_clk * world  # the N GPUs of the job
        gcups = tot_cells / (ms * 1e-3) / 1e9
        fill_gcups = tot_cells / (max(fill_all) * 1e-3) / 1e9 if max(fill_all) > 0 else None
        entries.append({"config": name, "algo": algo, "scoring": list(sc.astuple()), "pairs": int(total_pairs), "cells": tot_cells,
                        "gcups": gcups, "ms": ms, "kernel": r["kernel"], "ops_per_cell": W_OPS[algo],
                        "roofline_frac": gcups * 1e9 * W_OPS[algo] / p_int,
                        "fill_gcups": fill_gcups, "fill_roofline_frac": (fill_gcups * 1e9 * W_OPS[algo] / p_int) if fill_gcups else None,
                        "per_rank_ms": ms_all, "per_rank_cells": cells_all, "imbalance": ms / (sum(ms_all) / len(ms_all)),
                        "oracle_checked_pairs": int(sum(checked_all)), "oracle_checked_per_rank": [int(c) for c in checked_all],
                        "gpu_launches": int(launches), "scaling": "strong", "clocks": clocks, "note": note})

    # ---- configs[2]: Global + Local Gotoh, 10 M pairs of 250 bp, contiguous static shard ----
    n3 = args.config3_pairs
    lo, hi = shard.shard_range(n3, rank, world)
    seg = 2_500_000
    segs3 = [(a, min(seg, hi - a)) for a in range(lo, hi, seg)]
    for algo in ("ggotoh", "lgotoh"):
        entry("configs[2] %s 250 bp x %d" % ("GlobalGotohSA" if algo == "ggotoh" else "LocalGotohSA", n3), algo, S.affine(-3, -1, 1, -1),
              segs3, 0, 250, 250, 24,
              "ScoringSystem(-3,-1,1,-1); %d pairs split contiguously over %d rank(s), streamed through HBM in segments of <= %d pairs; "
              "score + traceback, inputs generated on the device" % (n3, world, seg), total_pairs=n3)

    # ---- configs[3]: Hirschberg / MyersMiller, 64 pairs of 100 kbp, whole pairs per rank ----
    n4, L4 = args.config4_pairs, 100_000
    lo, hi = shard.shard_range(n4, rank, world)
    gold = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "config4_100kbp.json")) as f:
            gold = json.load(f)
    except Exception:
        gold = None
    for algo, sc in (("hirschberg", S.linear(-1, 2, -1)), ("myersmiller", S.affine(-3, -1, 1, -1))):
        fp = None
        if gold and algo in gold["algos"] and gold["len"] == L4 and gold["seed"] == SEED and n4 <= gold["pairs"]:
            fp = gold["algos"][algo]["pairs"]
        entry("configs[3] %s 100 kbp x %d" % ("HirschbergSA" if algo == "hirschberg" else "MyersMillerSA", n4), algo, sc,
              [(lo, hi - lo)], 0, L4, L4, 0,
              "whole pairs per rank (%d rank(s)); every pair checked against the committed oracle fingerprints "
              "(tests/golden/config4_100kbp.json: score, ops_len, CRC-32 of the ops)" % world if fp else
              "whole pairs per rank (%d rank(s)); fingerprints file missing: unchecked here" % world, fingerprints=fp, total_pairs=n4)

    # ---- configs[4]: NW + SW over the mixed-length stream, shards balanced by sum(len1*len2), length-binned inside a device ----
    n5 = args.config5_pairs
    cuts, _ = _config5_cuts(SEED, n5, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    seg = 4_000_000
    segs5 = [(a, min(seg, hi - a)) for a in range(lo, hi, seg)]
    for algo, sc in (("nw", S.linear(-1, 2, -1)), ("sw", S.linear(-1, 1, -1))):
        entry("configs[4] %s mixed 50-1000 bp x %d" % ("NeedlemanWunschSA" if algo == "nw" else "SmithWatermanSA", n5), algo, sc,
              segs5, 1, 0, 0, 24,
              "independent U[50,1000] lengths; the stream is cut into %d contiguous shard(s) of equal sum(len1*len2); inside a device pairs are "
              "binned by shape (sorted by (ceil(len1/16), len2), 64 similar pairs per warp job, largest jobs first) and streamed through HBM "
              "in segments of <= %d pairs" % (world, seg), total_pairs=n5)
    return entries


def run_ours(args):
    import torch
    import torch.distributed as dist
    from seqalib_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # (NCCL_DEBUG is left to the caller: quiet_stdout() already keeps stdout to the one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lib = capi.Lib()  # raises if libseqa_cuda.so is missing
    n = args.pairs
    prm = capi.make_params("sw", gap=SCORING["gap"], match=SCORING["match"], mismatch=SCORING["mismatch"], allow=True,
                           device_first=local, device_count=1)
    stream = torch.cuda.Stream()  # the library launches on this stream; the timing events are recorded on it
    torch.cuda.set_stream(stream)
    ctx = capi.Ctx(lib, local, stream.cuda_stream)
    ctx.generate(prm, SEED, rank * n, n, 0, LEN, LEN)  # inputs resident in HBM, nothing crosses PCIe
    cells = ctx.cells()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        ctx.run()
    sampler.wait_first()
    ctx.sync()
    barrier()
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = sampler.mark()
    e0.record(stream)
    fill_ms, fill_launches = 0.0, 0
    for _ in range(args.steps):
        ctx.run()
    e1.record(stream)
    torch.cuda.synchronize()
    t1 = sampler.mark()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count() - l0
    fill_ms, fill_launches = ctx.last_fill_ms()  # CUDA events on the ctx stream around the last step's fill launches
    kernel = ctx.last_kernel()
    clocks = sampler.window(t0, t1)
    ctx.sync()
    ms_step = ms_total / args.steps
    total_cells = sum_over_ranks(float(cells))
    value = total_cells / (ms_step * 1e-3) / 1e9

    # ---- e2e: the C-ABI call with pinned host buffers, H2D + D2H inside the timed region ----
    tot_bases = n * 2 * LEN
    hb, ho1, ho2, hl1, hl2 = ctx.download_inputs(tot_bases)
    ctx.close()

    def pinned(shape, dt):
        t = torch.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=torch.uint8, pin_memory=True)
        return t.numpy().view(dt).reshape(shape)

    pb = pinned(tot_bases, np.uint8); pb[:] = hb
    po1 = pinned(n, np.uint64); po1[:] = ho1
    po2 = pinned(n, np.uint64); po2[:] = ho2
    pl1 = pinned(n, np.uint32); pl1[:] = hl1
    pl2 = pinned(n, np.uint32); pl2[:] = hl2
    res = capi.Results(n, tot_bases, pinned=pinned)
    e2e_steps = max(1, min(args.steps, 5))
    # bytes that really cross PCIe: the batch is dense and uniform, so the library sends the symbols only and derives
    # offsets and lengths on the device (SEQA_NO_DENSE_UPLOAD=1 sends all five arrays: + 24 bytes per pair)
    h2d = int(pb.nbytes) if not os.environ.get("SEQA_NO_DENSE_UPLOAD") else int(pb.nbytes + po1.nbytes + po2.nbytes + pl1.nbytes + pl2.nbytes)

    # the 2-bit input wire format (SEQA_FLAG_BASES_2BIT): what a caller that keeps its reads packed hands over -- 4 symbols
    # per byte, every sequence byte-aligned; packed here once, outside the timed region (it is the caller's storage format)
    packed_len = n * 2 * ((LEN + 3) // 4)
    pk = pinned(packed_len, np.uint8)
    _, pk1, pk2 = capi.pack_bases_2bit(pb, po1, po2, pl1, pl2, out=pk)
    pp1 = pinned(n, np.uint64); pp1[:] = pk1
    pp2 = pinned(n, np.uint64); pp2[:] = pk2

    def e2e_leg(flags):
        """the one-shot C-ABI call on pinned host buffers: H2D, kernels and D2H all inside the timed region"""
        p = capi.make_params("sw", gap=SCORING["gap"], match=SCORING["match"], mismatch=SCORING["mismatch"], allow=True,
                             device_first=local, device_count=1, flags=flags)
        ins = (pk, pp1, pp2, pl1, pl2) if flags & capi.FLAG_BASES_2BIT else (pb, po1, po2, pl1, pl2)
        for _ in range(2):
            lib.align_batch(p, *ins, res)
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            lib.align_batch(p, *ins, res)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        ms = max_over_ranks((w1 - w0) * 1e3 / e2e_steps)
        return ms, int(res.score.nbytes * 6 + res.ops_off.nbytes + int(res.c.ops_used)), int(res.score[:n].astype(np.int64).sum())

    # headline: both wire formats packed (2-bit symbols in, 2-bit ops out: what include/SequenceAlignment.h sends for ACGT
    # input); next to it the same call with the reference's own 8-bit symbols in, and with one byte per op out
    e2e_ms_b, d2h_b, ck_b = e2e_leg(0)
    e2e_ms_8, d2h_8, ck_8 = e2e_leg(capi.FLAG_OPS_2BIT)
    e2e_ms, d2h, ck_2 = e2e_leg(capi.FLAG_OPS_2BIT | capi.FLAG_BASES_2BIT)
    if not (ck_b == ck_8 == ck_2):
        raise SystemExit("bench.py: the wire formats disagree (score checksums %d / %d / %d)" % (ck_b, ck_8, ck_2))
    h2d_8 = h2d
    h2d = int(pk.nbytes) if not os.environ.get("SEQA_NO_DENSE_UPLOAD") else int(pk.nbytes + pp1.nbytes + pp2.nbytes + pl1.nbytes + pl2.nbytes)
    e2e_value = total_cells / (e2e_ms * 1e-3) / 1e9
    checksum = int(res.score[:n].astype(np.int64).sum())
    lib.L.seqa_cuda_trim()
    del pb, po1, po2, pl1, pl2, res, pk, pp1, pp2

    # ---- the template API itself (include/SequenceAlignment.h), rank 0 at N=1: compiled here, run as its own process ----
    api = None
    if world == 1 and not args.no_api:
        api = header_api_bench(n, local)

    # ---- the other BASELINE configs, sharded over the same ranks (strong scaling) ----
    configs = None
    if not args.no_configs:
        _, f_max, _ = measured_peaks()
        configs = run_secondary_configs(args, lib, capi, torch, dist, local, rank, world, stream, sampler, f_max)
    sampler.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, sm_max, how = measured_peaks()
    f_clk = (clocks.get("sm_mhz") or sm_max) * 1e6
    p_int = 148 * 128 * f_clk / 1e12  # T lane-ops/s at the SM clock observed during the run
    fill_s = fill_ms * 1e-3 / max(fill_launches, 1)
    cells_per_launch = cells / max(fill_launches, 1)
    achieved = cells_per_launch * W_OPS_PER_CELL / fill_s / 1e12
    trace_alg_bytes = cells_per_launch * ALG_TRACE_BITS / 8 + n * 2 * LEN / 4.0 / max(fill_launches, 1)
    # what the kernel really issues: ALU-pipe warp instructions per TWO cells, COUNTED from the SASS of the hot loop of
    # this build (tests/sass_count.py -> profiles/sass_counts.json), against the ALU-pipe ceiling measured by
    # tests/int_peak.py on this pool (63.8 lane-ops/clk/SM, profiles/r01_int_peak.json)
    alu_per_2cells, alu_src = sass_alu_per_two_cells(kernel)
    alu_ops = cells_per_launch / 2 * alu_per_2cells / fill_s / 1e12
    alu_peak = 148 * 63.8 * f_clk / 1e12
    step_achieved = value * 1e9 / max(world, 1) * W_OPS_PER_CELL / 1e12  # per GPU, whole step (prep + fill + walk + gather)
    roofline = {"bound": "int32_issue", "kernel": kernel, "achieved": achieved, "peak": p_int, "unit": "Tlane-op/s",
                "frac": achieved / p_int, "traffic": kernel_traffic(kernel, n), "ops_per_cell": W_OPS_PER_CELL,
                "frac_note": "achieved counts the W = 10 algorithmic int32 ops of a SW cell (SURVEY.md 8d); the kernel retires a "
                             "cell in ~%.2f issued ALU-pipe instructions (two cells per s16x2 instruction, fused add-max), so the "
                             "contract fraction can exceed 1; `alu_pipe.frac` is the hardware-bound figure" % (alu_per_2cells / 2),
                "step_frac": step_achieved / p_int,
                "step_note": "whole step per GPU (GCUPS x W / peak): the figure north_star's 'score+traceback >= 60 %' is about",
                "alu_pipe": {"achieved": alu_ops, "peak": alu_peak, "unit": "Tlane-op/s", "frac": alu_ops / alu_peak,
                             "instr_per_2_cells": alu_per_2cells, "source": alu_src,
                             "note": "issued ALU-pipe instructions vs the measured ALU-pipe rate (63.8 lane-ops/clk/SM)"},
                "kernel_ms_per_launch": fill_s * 1e3, "kernel_gcups": cells_per_launch / fill_s / 1e9,
                "peak_def": "148 SMs x 128 lane-ops/clk x SM clock observed under load (SURVEY.md 8d)",
                "hbm": {"bound": "hbm", "achieved": trace_alg_bytes / fill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": trace_alg_bytes / fill_s / 1e9 / hbm_peak, "peak_source": how,
                        "stored_bytes_per_cell": 0.5 * ((LEN + 15) // 16 * 16) * ((LEN + 3) // 4 * 4) / (LEN * LEN),
                        "stored_gbs": cells_per_launch * 0.5 * ((LEN + 15) // 16 * 16) * ((LEN + 3) // 4 * 4) / (LEN * LEN) / fill_s / 1e9}}

    cpu = None if not args.no_cpu else {"skipped": "--no-cpu"}
    if world > 1:
        cpu = {"skipped": "cpu_baseline is timed at N=1 only (the contract: rank 0 at N=1); see the --impl reference arm for every N"}
    if world == 1 and not args.no_cpu:
        threads = host_threads()
        sample_pairs = min(n, max(threads * 40000, 50000))  # ~5-15 s of CPU work on all host threads
        g, kind, sec = cpu_reference_gcups(sample_pairs, threads)
        cpu = {"value": g, "unit": "GCUPS", "cores": threads, "kind": kind, "seconds": sec,
               "sample": "first %d pairs of the same workload (same generator/seed)" % sample_pairs}

    line = {"metric": METRIC, "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2",
            "data": "synthetic",
            "config": {"workload": "SmithWatermanSA linear gap (-1,1,-1), %d random DNA pairs of %d bp per GPU, score+traceback" % (n, LEN),
                       "pairs_per_gpu": n, "len": LEN, "parallelism": "pairs sharded statically over %d GPU(s), no collective" % world, "numa_bound": numa is not None,
                       "l2": "no flush needed: every step streams %.1f GB of trace + %.0f MB of inputs through HBM (L2 is 126 MB)"
                             % (cells / 1e9 * 0.5 * 16 / 15 * 152 / 150, tot_bases / 1e6),
                       "result_checksum": checksum},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "api": "seqa_cuda_align_batch (pinned host buffers; SEQA_FLAG_BASES_2BIT | SEQA_FLAG_OPS_2BIT: symbols in and ops out "
                                                  "packed 4 per byte, the wire formats include/SequenceAlignment.h uses for ACGT input)",
                    "byte_bases": {"value": total_cells / (e2e_ms_8 * 1e-3) / 1e9, "ms_per_step": e2e_ms_8, "h2d_bytes_per_step": h2d_8,
                                   "d2h_bytes_per_step": d2h_8, "note": "same call with the reference's 8-bit symbols in (SEQA_FLAG_OPS_2BIT only: round 1's headline)"},
                    "byte_ops": {"value": total_cells / (e2e_ms_b * 1e-3) / 1e9, "ms_per_step": e2e_ms_b, "h2d_bytes_per_step": h2d_8,
                                 "d2h_bytes_per_step": d2h_b, "note": "8-bit symbols in, one byte per op out (flags = 0)"},
                    "api_packed": (api or {}).get("packed"), "api_list": (api or {}).get("list"),
                    "api_note": (api or {}).get("note", "template-API legs run at N=1 only")},
            "gpu_launches": int(launches), "clocks": clocks, "configs": configs}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-api", action="store_true", help="skip the C++ template-API legs (e2e.api_packed / e2e.api_list)")
    ap.add_argument("--no-configs", action="store_true", help="skip the secondary BASELINE configs (configs[2..4])")
    ap.add_argument("--config3-pairs", type=int, default=10_000_000, help="configs[2]: total Gotoh pairs of 250 bp (split over the ranks)")
    ap.add_argument("--config4-pairs", type=int, default=64, help="configs[3]: total pairs of 100 kbp (split over the ranks)")
    ap.add_argument("--config5-pairs", type=int, default=100_000_000, help="configs[4]: total mixed-length pairs (split over the ranks)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
