"""profiles/r02_traffic.json: DRAM bytes per launch of the kernel families bench.py reports rooflines for, taken from this
round's ncu captures (read here, no GPU):

    python tools/summarize_traffic.py profiles/r02_clip120_summary.json profiles/r02_metrics_launches.csv

The first argument is tools/summarize_ncu.py's summary of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` over tools/profile_call.py exact 120 (the 120-frame plan bench.py replays); the second the same metrics
over tools/bench_metrics.py (2048 fp32 pairs).  bench.py only copies these numbers into its JSON line."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    summ = json.load(open(sys.argv[1]))
    fam = {}
    for k in summ["kernels"]:
        name = k["kernel"].replace(" ", "")
        keys = []
        if name.startswith("gemm_tc2_kernel<0"):
            keys.append("uavsal_pw_gemm")
            if name.startswith("gemm_tc2_kernel<0,0,") and name.endswith(",2>"):
                keys.append("uavsal_pw_gemm/pair")
            if name.startswith("gemm_tc2_kernel<0,5,") and name.endswith(",2>"):
                keys.append("uavsal_pw_gemm/pair_q16")
        elif name.startswith("dw3x3"):
            keys.append("uavsal_dw3x3")
        elif name.startswith("dwproj_kernel"):
            keys.append("uavsal_dw_project")
        elif name.startswith("twa_step"):
            keys.append("uavsal_twa_step")
        for key in keys:
            f = fam.setdefault(key, {"dram_bytes": 0.0, "launches": 0, "us": 0.0})
            f["dram_bytes"] += (k["dram_read_MB"] + k["dram_write_MB"]) * 1e6
            f["launches"] += k["launches"]
            f["us"] += k["us"]
    out = {"source": os.path.relpath(sys.argv[1], ROOT), "kernels": {}}
    try:
        out["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        pass
    for key, f in fam.items():
        out["kernels"][key] = {"dram_bytes_per_launch": round(f["dram_bytes"] / f["launches"]), "launches": f["launches"],
                               "ncu_us_per_launch": round(f["us"] / f["launches"], 2)}
    out["plan_total"] = {"dram_GB": round(sum((k["dram_read_MB"] + k["dram_write_MB"]) for k in summ["kernels"]) / 1e3, 2),
                         "sum_kernel_ms": round(summ["total_us"] / 1e3, 3), "launches": summ["launches"], "frames": 120}
    if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
        rows = list(csv.DictReader(open(sys.argv[2])))
        rows = [r for r in rows if "metrics4" in r["kernel"]]
        if rows:
            r = rows[-1]
            pairs = int(os.environ.get("PAIRS", "2048"))
            alg = pairs * 3 * 360 * 640 * 4
            out["kernels"]["metrics4"] = {"kernel": r["kernel"], "dram_read_bytes": float(r["dram_read_bytes"]), "algorithmic_bytes": alg,
                                          "dram_read_ratio": round(float(r["dram_read_bytes"]) / alg, 3), "pairs": pairs,
                                          "ncu_us": float(r["duration_us"]), "source": os.path.relpath(sys.argv[2], ROOT)}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
