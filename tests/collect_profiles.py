"""Turns the scratch output of tests/gpu_final.sh (gpurun_out/<tag>_*) into the tracked evidence files under profiles/:
bench lines, `ncu --set full` summaries (tests/read_prof.py), the launch list with per-kernel shares, roofline.traffic
(profiles/<tag>_traffic.json: dram bytes per launch of every fill kernel captured), API phase timing, the pytest log.

    python tests/collect_profiles.py [tag=r02]
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def ncu_rows(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        return [], []
    return r[0], r[2:]


def summary(rep, dst, header):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "read_prof.py"), rep], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    with open(dst, "w") as f:
        f.write("# " + header + "\n" + out)


def launch_summary(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rr = csv.reader(lines)
    h = next(rr)
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for row in rr:
        if len(row) > vi:
            agg.setdefault(row[ki].split("(")[0], []).append(float(row[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    return ["%-44s n=%3d  avg %10.1f us   share %5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot) for k, v in agg.items()]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    for name in ("bench_1gpu.json", "bench_reference_arm.json", "bench_2gpu.json", "bench_4gpu.json", "bench_8gpu.json", "pytest_gpu.log", "api_timing.txt",
                 "launches_sw150_1M.csv", "smoke.log"):
        src = os.path.join(G, "%s_%s" % (tag, name))
        if os.path.exists(src) and os.path.getsize(src):
            shutil.copy(src, os.path.join(P, "%s_%s" % (tag, name)))
    ll = os.path.join(G, "%s_launches_sw150_1M.csv" % tag)
    if os.path.exists(ll):
        with open(os.path.join(P, "%s_launch_list_sw150_1M.txt" % tag), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --no-cpu --no-configs --no-api\n"
                    "# (resident steps of 1 M pairs + the waves of the e2e legs; per-launch times are cold-cache and serialised: read the SHARES)\n")
            f.write("\n".join(launch_summary(ll)) + "\n")
    traffic = {}
    for rep, dst, header in (("prof_pk", "ncu_pk_prep_fill_walk_sw150_1M.txt", "ncu --set full --clock-control none: bench.py headline (1 M x 150 bp SW): prep, fill, walk"),
                             ("prof_pkg", "ncu_pkg_fill_walk.txt", "ncu --set full: tests/bench_configs.py 'config3 GlobalGotoh' (200 k x 250 bp): packed affine fill (column codes) and walk"),
                             ("prof_ls_hb", "ncu_ls_sweep2_hb.txt", "ncu --set full: first ls_sweep2_kernel<0> launch (top level) of Hirschberg 64 x 100 kbp"),
                             ("prof_ls_mm", "ncu_ls_sweep2_mm.txt", "ncu --set full: first ls_sweep2_kernel<1> launch (top level) of MyersMiller 64 x 100 kbp")):
        path = os.path.join(G, "%s_%s.ncu-rep" % (tag, rep))
        if not os.path.exists(path):
            continue
        summary(path, os.path.join(P, "%s_%s" % (tag, dst)), header)
        hdr, rows = ncu_rows(path)
        if not hdr:
            continue
        kn, rd, wr, dur = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        units = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                               text=True).stdout.splitlines()))[1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for row in rows:
            name = row[kn]
            if "fill" not in name:
                continue
            b = float(row[rd]) * scale.get(units[rd], 1.0) + float(row[wr]) * scale.get(units[wr], 1.0)
            traffic[name] = {"dram_bytes_per_launch": b, "read": float(row[rd]) * scale.get(units[rd], 1.0), "write": float(row[wr]) * scale.get(units[wr], 1.0),
                             "duration": row[dur] + " " + units[dur], "capture": "%s_%s.ncu-rep" % (tag, rep)}
    # bench.py looks roofline.traffic up by ITS kernel name and scales by the pair count
    named = {}
    for name, t in traffic.items():
        if "pk_fill_kernel<1, 16, 2, 0" in name.replace("(bool)", "").replace("(int)", ""):
            named["pk_fill_sw_s16x2_t2"] = dict(t, pairs=1000000, ncu_kernel=name)
        if "pkg_fill_kernel<0, 16, 4" in name.replace("(bool)", "").replace("(int)", ""):
            named["pkg_fill_ggotoh_s16x2_t4"] = dict(t, pairs=200000, ncu_kernel=name)
    if named:
        with open(os.path.join(P, "%s_traffic.json" % tag), "w") as f:
            json.dump(named, f, indent=1, sort_keys=True)
            f.write("\n")
    print("collected:", sorted(x for x in os.listdir(P) if x.startswith(tag + "_")))


if __name__ == "__main__":
    main()
