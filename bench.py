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
    """DRAM bytes per launch of the dominant kernel from the committed `ncu` capture of this same command
    (profiles/r01_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum), scaled by the pair count."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)[kernel]
        return float(t["dram_bytes_per_launch"]) * pairs / float(t["pairs"])
    except Exception:
        return None


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
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for (t, line) in self.rows:
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


def cpu_reference_gcups(n_pairs, threads):
    """The reference's own CPU path (SmithWatermanSA::getAlignment with equal<char>, one aligner object per
    std::thread) on `n_pairs` pairs of the bench workload.  -> (gcups, kind, seconds)"""
    from oracle import pyoracle as orc
    from seqalib_b200 import synth
    if not os.path.exists(orc.ORACLE_SO):
        orc.build()
    bases, off1, off2, l1, l2 = synth.batch(SEED, 0, n_pairs, 0, LEN, LEN)
    sc = orc.Scoring.linear(SCORING["gap"], SCORING["match"], SCORING["mismatch"])
    cells = float((l1.astype(np.float64) * l2).sum())
    if orc.have_ref():
        sec, _ = orc.ref_bench("sw", sc, bases, off1, off2, l1, l2, threads)
        kind = "reference"
    else:
        sec, _ = orc.oracle_bench("sw", sc, bases, off1, off2, l1, l2, threads)
        kind = "port"
    return cells / sec / 1e9, kind, sec


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
    n = max(threads * 6000, 8000)  # ~1 s per step on all host threads at ~0.15 GCUPS/core
    vals, kind = [], "port"
    for k in range(args.warmup + args.steps):
        g, kind, sec = cpu_reference_gcups(n, threads)
        if k >= args.warmup:
            vals.append((g, sec))
    value = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    sample = "%d pairs of %d bp per step (same generator/seed as the GPU arm)" % (n, LEN)
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
        os.environ["NCCL_DEBUG"] = os.environ.get("SEQA_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
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
    clocks = sampler.stop(t0, t1)
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

    def e2e_leg(flags):
        """the one-shot C-ABI call on pinned host buffers: H2D, kernels and D2H all inside the timed region"""
        p = capi.make_params("sw", gap=SCORING["gap"], match=SCORING["match"], mismatch=SCORING["mismatch"], allow=True,
                             device_first=local, device_count=1, flags=flags)
        for _ in range(2):
            lib.align_batch(p, pb, po1, po2, pl1, pl2, res)
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            lib.align_batch(p, pb, po1, po2, pl1, pl2, res)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        ms = max_over_ranks((w1 - w0) * 1e3 / e2e_steps)
        return ms, int(res.score.nbytes * 6 + res.ops_off.nbytes + int(res.c.ops_used))

    # headline: the wire format the C++ host header requests (SEQA_FLAG_OPS_2BIT, 4 ops per byte); the one-byte-per-op
    # form of the same call is reported next to it
    e2e_ms_b, d2h_b = e2e_leg(0)
    e2e_ms, d2h = e2e_leg(capi.FLAG_OPS_2BIT)
    e2e_value = total_cells / (e2e_ms * 1e-3) / 1e9
    checksum = int(res.score[:n].astype(np.int64).sum())
    lib.L.seqa_cuda_trim()

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
    # what the kernel really issues: 5.25 ALU-pipe warp instructions per TWO cells (seqa_packed.cuh), against the
    # ALU-pipe ceiling measured by tests/int_peak.py on this pool (63.8 lane-ops/clk/SM, profiles/r01_int_peak.json)
    alu_ops = cells_per_launch / 2 * 5.25 / fill_s / 1e12
    alu_peak = 148 * 63.8 * f_clk / 1e12
    roofline = {"bound": "int32_issue", "kernel": kernel, "achieved": achieved, "peak": p_int, "unit": "Tlane-op/s",
                "frac": achieved / p_int, "traffic": kernel_traffic(kernel, n), "ops_per_cell": W_OPS_PER_CELL,
                "alu_pipe": {"achieved": alu_ops, "peak": alu_peak, "unit": "Tlane-op/s", "frac": alu_ops / alu_peak,
                             "note": "issued packed instructions (5.25 per 2 cells) vs the measured ALU-pipe rate"},
                "kernel_ms_per_launch": fill_s * 1e3, "kernel_gcups": cells_per_launch / fill_s / 1e9,
                "peak_def": "148 SMs x 128 lane-ops/clk x SM clock observed under load (SURVEY.md 8d)",
                "hbm": {"bound": "hbm", "achieved": trace_alg_bytes / fill_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": trace_alg_bytes / fill_s / 1e9 / hbm_peak, "peak_source": how,
                        "stored_bytes_per_cell": 0.5 * ((LEN + 15) // 16 * 16) * ((LEN + 3) // 4 * 4) / (LEN * LEN),
                        "stored_gbs": cells_per_launch * 0.5 * ((LEN + 15) // 16 * 16) * ((LEN + 3) // 4 * 4) / (LEN * LEN) / fill_s / 1e9}}

    cpu = None
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
                    "ms_per_step": e2e_ms, "api": "seqa_cuda_align_batch (pinned host buffers, SEQA_FLAG_OPS_2BIT: ops packed 4 per byte, the format include/SequenceAlignment.h requests)",
                    "byte_ops": {"value": total_cells / (e2e_ms_b * 1e-3) / 1e9, "ms_per_step": e2e_ms_b, "d2h_bytes_per_step": d2h_b,
                                 "note": "same call with one byte per op (flags = 0)"}},
            "gpu_launches": int(launches), "clocks": clocks}
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
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
