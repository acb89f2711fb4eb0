"""Per-instruction view of one kernel of an ncu report (--import-source on): stall samples, executed warp instructions,
average active threads.  usage: hot_sass.py report.ncu-rep kernel-regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], stdout=subprocess.PIPE, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
h = rows[0]
iS, iN, iE, iT, iA = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("Address")
data = []
for r in rows[1:]:
    if len(r) <= iT or not r[iA].startswith("0x"):
        break
    data.append((int(r[iA], 16), r[iS].strip(), int(r[iN] or 0), int(r[iE] or 0), int(r[iT] or 0)))
base = data[0][0]
totS = sum(d[2] for d in data); totE = sum(d[3] for d in data); totT = sum(d[4] for d in data)
print("instructions %d  samples %d  warp-instr executed %d  thread-instr %d  avg active threads %.1f" % (len(data), totS, totE, totT, totT / max(totE, 1)))
# cumulative by address ranges of 16 instructions
print("-- by block of 16 instructions: offset, samples%, exec%, avg threads")
for k in range(0, len(data), 16):
    blk = data[k:k + 16]
    s = sum(d[2] for d in blk); e = sum(d[3] for d in blk); t = sum(d[4] for d in blk)
    if s * 100.0 / totS >= 1.0 or e * 100.0 / totE >= 1.5:
        print("  0x%04x  %5.1f%%  %5.1f%%  %5.1f   %s" % (blk[0][0] - base, s * 100.0 / totS, e * 100.0 / totE, t / max(e, 1), blk[0][1][:50]))
print("-- top instructions by samples")
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print("  0x%04x  %5.2f%%  exec %9d  thr %4.1f  %s" % (d[0] - base, d[2] * 100.0 / totS, d[3], d[4] / max(d[3], 1), d[1][:70]))
