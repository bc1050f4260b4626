"""Summarises an `ncu --set full --import-source on` report: headline metrics + warp-stall reasons per kernel and the
SASS instructions with the most stall samples.   python tools/ncu_stalls.py report.ncu-rep [topN]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h = r[0]
    want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg",
            "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum", "launch__grid_size"]
    for row in r[2:]:
        print("===", row[h.index("Kernel Name")][:90])
        for w in want:
            if w in h:
                print(f"   {w:70s} {row[h.index(w)]} {r[1][h.index(w)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = [i for i, x in enumerate(rows) if x and x[0] == "Address"]
    names = [rows[i - 1][1] if i > 0 and len(rows[i - 1]) > 1 else "?" for i in hdr]
    for bi, start in enumerate(hdr):
        end = hdr[bi + 1] - 1 if bi + 1 < len(hdr) else len(rows)
        hh = rows[start]
        ci = {n: i for i, n in enumerate(hh)}
        data = [x for x in rows[start + 1:end] if len(x) >= len(hh)]
        stalls = [n for n in hh if n.startswith("stall_") and "Not Issued" not in n]
        tot = {s: 0 for s in stalls}
        for x in data:
            for s in stalls:
                if x[ci[s]].isdigit():
                    tot[s] += int(x[ci[s]])
        S = sum(tot.values()) or 1
        print(f"\n=== stalls: {names[bi][:90]}  (samples {S})")
        print("   " + ", ".join(f"{s[6:]} {100 * v / S:.0f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        top = sorted((x for x in data if x[ci["# Samples"]].isdigit()), key=lambda x: -int(x[ci["# Samples"]]))[:topn]
        for x in top:
            st = sorted(((s[6:], int(x[ci[s]])) for s in stalls if x[ci[s]].isdigit() and int(x[ci[s]]) > 0),
                        key=lambda z: -z[1])[:2]
            print(f"   {x[ci['# Samples']]:>6s} {x[ci['Instructions Executed']]:>9s}  {x[ci['Source']][:72]:72s} {st}")


if __name__ == "__main__":
    main()
