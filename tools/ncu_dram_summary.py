"""profiles/r2_gemm_dram.json from an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
-k regex:gemm_f16` capture of one training step (bench.py --profile-step):
    python tools/ncu_dram_summary.py gpurun_out/r2_gemm_dram.csv [algorithmic_bytes_json] > profiles/r2_gemm_dram.json"""
import collections
import csv
import json
import sys


def main():
    with open(sys.argv[1]) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.defaultdict(dict)
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        elif u in ("ns", "nsecond"):
            v *= 1e-3
        elif u in ("ms", "msecond"):
            v *= 1e3
        per[x["ID"]][x["Metric Name"]] = v
        per[x["ID"]]["name"] = x["Kernel Name"]
    n = len(per)
    rd = sum(p.get("dram__bytes_read.sum", 0.0) for p in per.values())
    wr = sum(p.get("dram__bytes_write.sum", 0.0) for p in per.values())
    us = sum(p.get("gpu__time_duration.sum", 0.0) for p in per.values())
    out = {"launches": n, "dram_bytes_read_per_launch": round(rd / n), "dram_bytes_write_per_launch": round(wr / n),
           "dram_bytes_per_launch": round((rd + wr) / n), "avg_us_per_launch_under_ncu": round(us / n, 2),
           "dram_gbs_under_ncu": round((rd + wr) / (us * 1e-6) / 1e9, 1) if us else None,
           "source": sys.argv[1], "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
           "--clock-control none -k regex:gemm_f16 over one eager training step (bench.py --profile-step)"}
    if len(sys.argv) > 2:
        out.update(json.load(open(sys.argv[2])))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
