"""Aggregates `ncu --page source --csv --print-source cuda,sass`: executed instructions and stall samples per CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
files = {}; cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[2] != "-": continue          # SASS rows have an address; source rows have "-"
    try:
        line = int(r[0]); n = int(r[7] or 0); s = int(r[6] or 0)
    except ValueError:
        continue
    out.append((n, s, cur, line, r[1].strip()[:110]))
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total executed %d, samples %d" % (tot, tots))
for n, s, f, line, src in sorted(out, reverse=True)[:top]:
    print("%5.2f%% ex %5.2f%% smp  %s:%d  %s" % (100.0 * n / max(tot, 1), 100.0 * s / max(tots, 1), f, line, src))
