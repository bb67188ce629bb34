"""Aggregate an ncu SASS profile by SOURCE LINE: joins `ncu --page source --csv --print-source sass` (samples and executed
instructions per SASS address) with the line table of the cubin (`nvdisasm -g`).

usage: python scripts/ncu_by_line.py <report.ncu-rep> <kernel symbol regex> [library.so] [top]
Prints, per source file:line of the kernel: warp instructions executed, stall samples, shared wavefronts; and the same
summed over named line ranges (functions) when scripts/ncu_regions.json maps the file to ranges."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, sym_re = sys.argv[1], re.compile(sys.argv[2])
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "marbler_b200", "libmarbler_b200.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

primary = os.environ.get("NCU_PRIMARY")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin") or "sm_100" not in f:
        continue
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
    cur_fun, cur_line, inside, last_primary = None, None, False, None
    for l in dis:
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            cur_fun = m.group(1)
            inside = bool(sym_re.search(cur_fun))
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            # the innermost frame comes first; "inlined at" annotations of the callers follow on the same line
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            # NCU_PRIMARY=<file>: instructions inlined from other files (intrinsics headers, common.cuh) are charged to
            # the last line of the primary file seen before them (nvdisasm prints only the innermost frame)
            if primary and cur_line[0] != primary and last_primary:
                cur_line = last_primary
            elif primary and cur_line[0] == primary:
                last_primary = cur_line
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", l)
        if m and cur_line:
            line_of[int(m.group(1), 16)] = cur_line
    if line_of:
        break

csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(csv_txt.split("\n")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(rows[hi])}
agg = collections.defaultdict(lambda: [0, 0, 0, 0])      # instr, samples, wavefronts, ideal wavefronts
base = None
for r in rows[hi + 1:]:
    if len(r) < len(ix) or r[0] == "Address":
        continue
    addr = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    if base is None:
        base = addr
    key = line_of.get(addr - base, ("?", 0))

    def num(c):
        try:
            return int(r[ix[c]] or 0)
        except (ValueError, KeyError):
            return 0
    a = agg[key]
    a[0] += num("Instructions Executed"); a[1] += num("# Samples")
    a[2] += num("L1 Wavefronts Shared"); a[3] += num("L1 Wavefronts Shared Ideal")
ti, ts, tw = (sum(a[k] for a in agg.values()) for k in (0, 1, 2))
print("total: %d warp instructions, %d samples, %d shared wavefronts" % (ti, ts, tw))
print("%-28s %12s %7s %8s %7s %10s" % ("file:line", "instr", "%", "samples", "%", "wavefronts"))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-28s %12d %6.1f%% %8d %6.1f%% %10d" % ("%s:%d" % key, a[0], 100.0 * a[0] / max(ti, 1), a[1], 100.0 * a[1] / max(ts, 1), a[2]))
# by contiguous ranges given on stdin-free env var: NCU_REGIONS="name:file:lo-hi,..."
regions = os.environ.get("NCU_REGIONS")
if regions:
    print()
    for spec in regions.split(","):
        name, f, rng = spec.split(":")
        lo, hi_ = (int(v) for v in rng.split("-"))
        s = [0, 0, 0, 0]
        for (ff, ln), a in agg.items():
            if ff == f and lo <= ln <= hi_:
                for k in range(4):
                    s[k] += a[k]
        print("%-24s instr %6.1f%%  samples %6.1f%%  wavefronts %6.1f%% (%.2fx ideal)" % (
            name, 100.0 * s[0] / max(ti, 1), 100.0 * s[1] / max(ts, 1), 100.0 * s[2] / max(tw, 1), s[2] / max(s[3], 1)))
