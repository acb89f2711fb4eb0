"""Counts the instruction mix of the hot loops of the built library from its SASS (cuobjdump, no GPU needed) so that the
"ALU-pipe instructions per two cells" figure bench.py reports is DERIVED from the committed build, not typed in.

    python tests/sass_count.py            # rewrites profiles/sass_counts.json and profiles/r02_sass_<kernel>.txt excerpts

For every fill kernel: the hot loop is the innermost backward-branch loop that holds the most VIADDMNMX instructions
(one iteration = one column group = 4 columns x 16 rows x 2 pairs = 128 cells).  Instruction classes:
  alu   -- ALU-pipe-only instructions (measured 63.8 lane-ops/clk/SM on B200, profiles/r01_int_peak.json):
           VIADDMNMX*, VIADD*, VIMNMX*, VIMNMX3*, PRMT, LOP3, SHF, SEL, ISETP, IABS, BMSK, SGXT, LEA (non-IMAD integer forms)
  fma   -- integer ops that also issue on the FMA pipes: IMAD*, IADD3 (measured 127.5)
  mem   -- LDG / STG / LDS / STS / LD / ST / ATOM / RED / CCTL
  other -- MOV / control / uniform-datapath instructions
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "seqalib_b200", "libseqa_cuda.so")

# bench.py's kernel names -> mangled functions, cells per hot-loop iteration
KERNELS = {
    "pk_fill_sw_s16x2_t2": ("_Z14pk_fill_kernelILb1ELi16ELi2ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_nw_s16x2_t2": ("_Z14pk_fill_kernelILb0ELi16ELi2ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_sw_s16x2_t4": ("_Z14pk_fill_kernelILb1ELi16ELi4ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_nw_s16x2_t4": ("_Z14pk_fill_kernelILb0ELi16ELi4ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_sw_s16x2_t8": ("_Z14pk_fill_kernelILb1ELi16ELi8ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_nw_s16x2_t8": ("_Z14pk_fill_kernelILb0ELi16ELi8ELb0ELb0EEv6PkArgs", 128),
    "pk_fill_sw_s16x2_t4_gb": ("_Z14pk_fill_kernelILb1ELi16ELi4ELb1ELb0EEv6PkArgs", 128),
    "pk_fill_nw_s16x2_t4_gb": ("_Z14pk_fill_kernelILb0ELi16ELi4ELb1ELb0EEv6PkArgs", 128),
    "pkg_fill_ggotoh_s16x2_t4": ("_Z15pkg_fill_kernelILb0ELi16ELi4ELb1EEv6PkArgs", 128),
    "pkg_fill_lgotoh_s16x2_t4": ("_Z15pkg_fill_kernelILb1ELi16ELi4ELb1EEv6PkArgs", 128),
}

ALU = ("VIADDMNMX", "VIADD", "VIMNMX", "VIMNMX3", "PRMT", "LOP3", "SHF", "SEL", "ISETP", "IABS", "BMSK", "SGXT", "LEA", "PLOP3", "POPC", "FLO")
FMA = ("IMAD", "IADD3", "IADD")
MEM = ("LDG", "STG", "LDS", "STS", "LD", "ST", "ATOM", "ATOMG", "RED", "CCTL", "LDC", "LDCU", "LDSM", "STSM")


def classify(op):
    base = op.split(".")[0]
    if base in ALU:
        return "alu"
    if base in FMA:
        return "fma"
    if base in MEM:
        return "mem"
    return "other"


def disassemble(fn):
    txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, SO], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    ins = []
    for line in txt.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        body = m.group(2).strip()
        pred = re.match(r"^@!?U?P\d+\s+", body) is not None
        body = re.sub(r"^@!?U?P\d+\s+", "", body)
        op = body.split()[0] if body else "?"
        tgt = None
        mb = re.match(r"BRA(?:\.\w+)*\s+(!?U?P\d+,\s*)?(0x[0-9a-f]+)", body)
        if mb:
            tgt = int(mb.group(2), 16)
            pred = pred or mb.group(1) is not None  # BRA.U !UP0, target: conditional through its operand
        ins.append((addr, op, body, tgt, line, pred))
    return ins


def executed_path(loop):
    """One iteration's instructions: the loop body is a DAG of forward branches (the rare corner-capture variant of the
    global aligners -- taken twice per pair -- and the guarded loads of the next column group).  Among all paths from
    the loop head to the back edge: run the recurrence ONCE (fewest VIADDMNMX), take the guarded loads (most memory
    instructions), and of those the SHORTEST path (the variant without the capture code)."""
    n = len(loop)
    index = {x[0]: i for i, x in enumerate(loop)}
    memo = {}

    def weight(i):
        op = loop[i][1]
        return (1 if op.startswith("VIADDMNMX") else 0, 1 if classify(op) == "mem" else 0)

    def walk(i):  # -> {(viaddmnmx, mem): (length, [indices])}: shortest path per key from i to the end of the loop
        if i >= n - 1:
            return {weight(n - 1): (1, [n - 1])}
        if i in memo:
            return memo[i]
        addr, op, body, tgt, line, pred = loop[i]
        succ = []
        if op.startswith("BRA") and tgt is not None and tgt in index and index[tgt] > i:
            succ.append(index[tgt])
            if pred:
                succ.append(i + 1)
        else:
            succ.append(i + 1)
        wv, wm = weight(i)
        res = {}
        for sx in succ:
            for (cv, cm), (ln, path) in walk(sx).items():
                key = (cv + wv, cm + wm)
                if key not in res or res[key][0] > ln + 1:
                    res[key] = (ln + 1, [i] + path)
        memo[i] = res
        return res
    sys.setrecursionlimit(100000)
    paths = walk(0)
    need_v = min(k[0] for k in paths if k[0] > 0)
    need_m = max(k[1] for k in paths if k[0] == need_v)
    return [loop[i] for i in paths[(need_v, need_m)][1]]


def hot_loop(ins):
    """innermost loop (backward BRA) holding the most VIADDMNMX"""
    best = None
    for k, (addr, op, body, tgt, _, _p) in enumerate(ins):
        if tgt is None or tgt > addr:
            continue
        lo = next(i for i, x in enumerate(ins) if x[0] == tgt)
        body_ins = ins[lo:k + 1]
        inner = any(x[3] is not None and x[3] <= x[0] and x[3] > tgt for x in body_ins[:-1])  # holds another backward branch
        n = sum(1 for x in body_ins if x[1].startswith("VIADDMNMX"))
        if inner or n == 0:
            continue
        if best is None or n > best[0]:
            best = (n, lo, k)
    return best


def main():
    out = {}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for name, (fn, cells) in KERNELS.items():
        ins = disassemble(fn)
        if not ins:
            continue
        hl = hot_loop(ins)
        if not hl:
            continue
        _, lo, hi = hl
        static_loop = ins[lo:hi + 1]
        loop = executed_path(static_loop)
        mix, classes = {}, {"alu": 0, "fma": 0, "mem": 0, "other": 0}
        for (_, op, _, _, _, _) in loop:
            mix[op] = mix.get(op, 0) + 1
            classes[classify(op)] += 1
        out[name] = {"function": fn, "loop": "0x%04x..0x%04x" % (static_loop[0][0], static_loop[-1][0]), "instructions": len(loop),
                     "static_instructions_in_loop": len(static_loop),
                     "cells_per_iteration": cells, "classes": classes,
                     "alu_per_2_cells": classes["alu"] * 2.0 / cells, "issued_per_cell": len(loop) / float(cells),
                     "mix": dict(sorted(mix.items(), key=lambda kv: -kv[1]))}
        if name in ("pk_fill_sw_s16x2_t2", "pk_fill_sw_s16x2_t4", "pkg_fill_ggotoh_s16x2_t4"):
            with open(os.path.join(ROOT, "profiles", "r02_sass_%s.txt" % name), "w") as f:
                f.write("# hot loop of %s (%s): one iteration = one 4-column group x 16 rows x 2 pairs = %d cells\n" % (name, fn, cells))
                f.write("# %d instructions; classes %r; ALU-pipe instructions per 2 cells = %.3f\n" % (len(loop), classes, classes["alu"] * 2.0 / cells))
                f.write("# mix: %s\n" % ", ".join("%s x%d" % kv for kv in sorted(mix.items(), key=lambda kv: -kv[1])))
                for x in loop:
                    f.write(x[4].rstrip() + "\n")
    with open(os.path.join(ROOT, "profiles", "sass_counts.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    for k, v in out.items():
        sys.stdout.write("%-28s %4d instr / %d cells: alu %d fma %d mem %d other %d -> %.3f ALU per 2 cells\n" % (
            k, v["instructions"], v["cells_per_iteration"], v["classes"]["alu"], v["classes"]["fma"], v["classes"]["mem"],
            v["classes"]["other"], v["alu_per_2_cells"]))


if __name__ == "__main__":
    main()
