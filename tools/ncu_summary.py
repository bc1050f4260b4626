"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).
    python tools/ncu_summary.py gpurun_out/launches.csv [steps_in_capture] > profiles/rNN_launches_summary.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = re.sub(r"\(.*", "", re.sub(r"^void ", "", x["Kernel Name"]))
        name = re.sub(r"<.*", "", name) if not name.startswith("lav::") else name
        tot[name][0] += 1
        tot[name][1] += v
    S = sum(v[1] for v in tot.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"capture = {steps} training step(s) (bench.py --quick); per-launch times are cold-cache and serialised, "
          f"compare SHARES.\n")
    print(f"total kernel time {S / 1e3:.2f} ms = {S / 1e3 / steps:.2f} ms/step, "
          f"{sum(v[0] for v in tot.values())} launches = {sum(v[0] for v in tot.values()) // steps}/step\n")
    print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"| `{n[:90]}` | {c / steps:.0f} | {t / 1e3 / steps:.3f} | {100 * t / S:.1f}% |")


if __name__ == "__main__":
    main()
