"""Summarise an `ncu --page source --csv --print-source sass` dump: opcode mix, stall reasons, hot source lines."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops, stat, stalls = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    try:
        n = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    ops[op] += n; stat[op] += 1; tot += n
    for s in stall_cols:
        v = r[ix[s]]
        if v:
            stalls[s] += int(v)
print("total warp instr", tot, "static", sum(stat.values()))
for op, n in ops.most_common(22):
    print("%-10s exec %10d (%5.1f%%) static %d" % (op, n, 100.0 * n / max(tot, 1), stat[op]))
ts = sum(stalls.values())
for s, n in stalls.most_common(10):
    print("%-28s %5.1f%%" % (s, 100.0 * n / max(ts, 1)))
