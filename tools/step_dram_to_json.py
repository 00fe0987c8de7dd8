"""Turn the long-format CSV of `ncu --metrics ... --csv` (tools/gpu_profile_r2.sh: per-launch duration, DRAM bytes, tensor-pipe
active %, L2 hit rate over one bench.py step) into the per-launch JSON that bench.py's `roofline.traffic` reads.
Usage: python tools/step_dram_to_json.py gpurun_out/step_dram.csv profiles/r02_step_per_launch_dram.json"""
import csv
import json
import re
import sys


def main(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows, order = {}, []
    for r in csv.DictReader(lines):
        i = int(r["ID"])
        if i not in rows:
            name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("hk::", "")
            rows[i] = {"kernel": name, "grid": r["Grid Size"], "block": r["Block Size"]}
            order.append(i)
        rows[i][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    out = [rows[i] for i in order]
    with open(dst, "w") as f:
        json.dump(out, f, indent=0)
    tot = sum(r.get("gpu__time_duration.sum", 0.0) for r in out)
    print(f"{dst}: {len(out)} launches, {tot / 1e6:.3f} ms")
    agg = {}
    for r in out:
        a = agg.setdefault(r["kernel"], [0, 0.0, 0.0])
        a[0] += 1; a[1] += r.get("gpu__time_duration.sum", 0.0)
        a[2] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:60s} x{n:3d} {t / 1e6:8.3f} ms {b / 1e6:10.1f} MB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
