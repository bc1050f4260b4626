"""Per-family and per-(kernel, grid) breakdown of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    by, fam = collections.defaultdict(list), collections.defaultdict(float)
    for x in rows:
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = re.sub(r"\(.*", "", re.sub(r"^void ", "", x["Kernel Name"]))
        by[(name, x["Grid Size"])].append(v)
        f = "gemm" if "gemm" in name else "attention" if "attn" in name else re.sub(r"<.*", "", name)
        fam[f] += v
    tot = sum(fam.values())
    print(f"total {tot / 1e3:.2f} ms, {len(rows)} launches")
    for f, v in sorted(fam.items(), key=lambda kv: -kv[1])[:14]:
        print(f"  {f:45s} {v / 1e3:7.3f} ms {100 * v / tot:5.1f}%")
    out = sorted(((sum(v), len(v), n, g, min(v), max(v)) for (n, g), v in by.items()), reverse=True)
    for t, c, n, g, mn, mx in out[:top]:
        print(f"{t / 1e3:7.3f} ms  n={c:3d} avg={t / c:7.1f}us min={mn:6.1f} max={mx:6.1f} grid={g:>16s} {n[:58]}")


if __name__ == "__main__":
    main()
