"""Runs the GPU test-suite on the B200 box with per-test isolation where needed.

    python tools/gpu_tests.py [pytest args / test paths]

1. one pytest process over everything (fast path), per-test timeout;
2. if anything failed, every failed test id is re-run alone in a fresh process (a trapped kernel poisons
   the CUDA context of its process, so cascaded failures are separated from real ones).
Logs go to gpurun_out/pytest_*.log.
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def run(args, log, timeout):
    with open(log, "w") as f:
        try:
            r = subprocess.run([sys.executable, "-m", "pytest"] + args, cwd=ROOT, stdout=f, stderr=subprocess.STDOUT,
                               timeout=timeout)
            return r.returncode
        except subprocess.TimeoutExpired:
            f.write("\n[gpu_tests] TIMEOUT\n")
            return 124


def main():
    os.makedirs(OUT, exist_ok=True)
    targets = sys.argv[1:] or ["tests"]
    log = os.path.join(OUT, "pytest_all.log")
    rc = run(["-m", "gpu", "-q", "-rfE", "--timeout=240", "-p", "no:cacheprovider"] + targets, log, 1500)
    text = open(log).read()
    print(text[-3000:])
    if rc == 0:
        print("[gpu_tests] ALL PASSED")
        return 0
    failed = re.findall(r"^(?:FAILED|ERROR) (\S+)", text, flags=re.M)
    failed = list(dict.fromkeys(failed))[:40]
    print(f"[gpu_tests] {len(failed)} failed in the shared process; re-running each in isolation")
    real = []
    for i, nodeid in enumerate(failed):
        l = os.path.join(OUT, f"pytest_iso_{i:02d}.log")
        r = run(["-q", "-x", "--timeout=240", "-p", "no:cacheprovider", nodeid], l, 400)
        tail = open(l).read()[-1500:]
        status = "PASS" if r == 0 else "FAIL"
        print(f"--- [{status}] {nodeid}")
        if r != 0:
            real.append(nodeid)
            print(tail)
    print(f"[gpu_tests] isolated failures: {len(real)} / {len(failed)}")
    for n in real:
        print("   ", n)
    return 1 if real else 0


if __name__ == "__main__":
    sys.exit(main())
