"""Print selected raw metrics (substring filters) of each launch in an ncu report.
  python tools/ncu_metrics.py report.ncu-rep [kernel-substring] -- filter1 filter2 ..."""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    args = sys.argv[2:]
    ksub = ""
    if args and args[0] != "--":
        ksub = args[0]
        args = args[1:]
    filters = [a for a in args if a != "--"] or ["warp_issue_stalled", "pipe_alu.avg.pct", "pipe_fma.avg.pct", "pipe_lsu", "issue_active.avg.pct",
                                                 "lts__throughput.avg", "l1tex__throughput.avg", "lts__t_bytes.sum ", "tensor_cycles_active.avg.pct"]
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        if ksub not in name:
            continue
        print("==", name[:70])
        for i, m in enumerate(h):
            if any(f in m for f in filters) and r[i] not in ("", "0", "0.000000"):
                print(f"   {m:90s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
