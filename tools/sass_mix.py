"""Aggregates an `ncu --page source --csv` dump: executed warp-instructions and stall samples by SASS opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter(); st = collections.Counter(); tot = 0; tots = 0
stalls = collections.Counter()
scols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    ex[op] += n; st[op] += s; tot += n; tots += s
    for h in scols:
        stalls[h] += int(r[ix[h]] or 0)
print("static instructions: %d, executed warp-instructions: %d, samples: %d" % (len(rows) - 2, tot, tots))
for op, n in ex.most_common(25):
    print("%-10s exec %6.2f%%  samples %6.2f%%" % (op, 100.0 * n / tot, 100.0 * st[op] / max(tots, 1)))
print("stall samples:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(tots, 1)) for h, v in stalls.most_common(10)))
