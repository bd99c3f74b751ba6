"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python tools/launch_summary.py gpurun_out/launches_r02.csv "<command that was profiled>" > profiles/r02_launches_by_kernel.txt
"""
import csv
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = name.replace("void ", "").replace("sb::", "").replace("(anonymous namespace)::", "<unnamed>::")
    return name.split("(CUtensorMap")[0].split("(const ")[0].split("(sb::")[0][:70]


def main(path, cmd):
    lines = [l for l in open(path, newline="") if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
        k = short(r["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    total = sum(tot.values())
    print(f"# launch list of `{cmd}`")
    print("# under `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are serialised and cold-cache:")
    print("# the SHARE of each kernel is what compares with the live CUDA-event profile of bench.py).")
    print(f"# {sum(cnt.values())} launches captured (the capture stops at ncu's -c limit; torch's weight-initialisation kernels are the at:: rows)\n")
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>9s}")
    for k in sorted(tot, key=lambda k: -tot[k])[:40]:
        print(f"{k:70s} {cnt[k]:8d} {tot[k]:12.1f} {100 * tot[k] / total:6.2f}% {tot[k] / cnt[k]:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
