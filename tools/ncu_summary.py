"""Summarises one `ncu --set full` report (read here, no GPU): the launch metrics the roofline discussion uses, the SASS opcode mix and
the stall reasons; optionally records the measured DRAM traffic in profiles/traffic.json for bench.py.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.md> "<title>" [--traffic <config> <envs>]"""
import collections, csv, hashlib, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, out, title = sys.argv[1:4]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = raw[0], raw[1]
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
lines = ["# " + title, ""]
for li, vals in enumerate(raw[2:]):
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    lines += ["## launch %d: `%s`" % (li, d.get("Kernel Name", "?")), "", "| metric | unit | value |", "|---|---|---|"]
    for h in WANT:
        if h in d: lines.append("| %s | %s | %s |" % (h, u[h], d[h]))
    st = sorted(((float(d[h].replace(",", "")), h) for h in STALL if d.get(h)), reverse=True)[:8]
    lines += ["", "Stall cycles per issued instruction: " + ", ".join("%s %.2f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v) for v, h in st), ""]
    if li == 0 and "--traffic" in sys.argv:
        i = sys.argv.index("--traffic"); cfg, envs = sys.argv[i + 1], int(sys.argv[i + 2])
        def num(x, un): return float(x.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un]
        tr = num(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        h = hashlib.sha256()
        for f in ("rkfd_core.cuh", "rkfd_kernel.cuh", "rkfd_math.cuh", "rkfd_types.h", "rkfd_volume.cuh", "rkfd_kernel_variant.cu"):
            h.update(open(os.path.join(ROOT, "roki-fd_b200", "csrc", f), "rb").read())
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        tj = json.load(open(tp)) if os.path.exists(tp) else {}
        if "dram_bytes_per_launch" in tj: tj = {}          # round-1 format
        tj[cfg] = {"dram_bytes_per_launch": tr, "envs": envs, "kernel": d.get("Kernel Name"), "capture": os.path.relpath(out, ROOT), "kernel_source_hash": h.hexdigest()[:16]}
        json.dump(tj, open(tp, "w"), indent=1)
        lines += ["DRAM traffic per launch %.1f MB = %.0f B per env-step (recorded in profiles/traffic.json for `%s`, %d envs)." % (tr / 1e6, tr / envs, cfg, envs), ""]
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
if len(src) > 2:
    h = src[1]; ix = {x: i for i, x in enumerate(h)}
    ex, sm = collections.Counter(), collections.Counter(); tot = tots = 0
    for r in src[2:]:
        if len(r) < len(h): continue
        t = r[ix["Source"]].split()
        if not t: continue
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        n, s = int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)
        ex[op] += n; sm[op] += s; tot += n; tots += s
    f64 = sum(ex[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    lines += ["## SASS mix of launch 0 (%d static instructions, %d executed warp-instructions, %.1f %% fp64)" % (len(src) - 2, tot, 100.0 * f64 / max(tot, 1)), "",
              "| opcode | executed | stall samples |", "|---|---|---|"]
    for op, n in ex.most_common(16): lines.append("| %s | %.2f %% | %.2f %% |" % (op, 100.0 * n / tot, 100.0 * sm[op] / max(tots, 1)))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
