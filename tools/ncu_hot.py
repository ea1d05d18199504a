"""Summarise an .ncu-rep (read here, no GPU): key counters of every captured launch + the top stall instructions of the first.
usage: python tools/ncu_hot.py file.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__cluster_max_active", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "lts__t_bytes.sum", "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:100])
    for k in hdr:
        if any(k.startswith(x) for x in KEYS) and "per_second" not in k and ".max" not in k and ".min" not in k:
            print("  %-80s %s %s" % (k, d[k], rows[1][hdr.index(k)]))
    st = sorted(((float(d[k]), k) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)
    print("  stalls/issue:", ", ".join("%s %.2f" % (k.split("stalled_")[1].split("_per_")[0], v) for v, k in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
col, srcc, ex = ci["Warp Stall Sampling (All Samples)"], ci["Source"], ci["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []; tot = 0
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name": break
        continue
    try: v = int(r[col])
    except ValueError: continue
    tot += v
    top = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    out.append((v, idx, r[srcc].strip()[:70], r[ex], top))
print("total samples", tot, "instructions", len(out))
blk = 64
print("samples per %d-instruction block:" % blk, [(b, sum(v for v, *_ in out[b:b + blk])) for b in range(0, len(out), blk) if sum(v for v, *_ in out[b:b + blk]) > tot / 100])
for v, idx, s, e, top in sorted(out, reverse=True)[:topn]:
    print("%6d %5.1f%% #%4d %10s %-70s %s" % (v, 100.0 * v / max(tot, 1), idx, e, s, top))
