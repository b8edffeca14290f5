"""Top source lines by warp-stall samples for one kernel of an ncu report (needs -lineinfo + --import-source on).
  python tools/ncu_source_hot.py report.ncu-rep <kernel-regex> [launch-skip] [top]"""
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(float(x or 0))
    except ValueError:
        return 0


def main():
    path, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}",
                          "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = None
    fname = ""
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            print("kernel:", r[1][:100])
        elif r[0] == "Line No":
            h = r
        elif h and r[0] not in ("",):
            lines.append((fname, r))
    si = h.index("# Samples")
    ie = h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    total = sum(num(r[si]) for _, r in lines)
    tot_inst = sum(num(r[ie]) for _, r in lines)
    print(f"total samples {total}   warp instructions {tot_inst}")
    for f, r in sorted(lines, key=lambda fr: -num(fr[1][si]))[:top]:
        n = num(r[si])
        st = sorted(((num(r[c]), h[c][6:]) for c in stall_cols), reverse=True)[:3]
        why = ", ".join(f"{k}={v}" for v, k in st if v)
        print(f"{n:7d} {100.0 * n / max(total, 1):5.1f}%  inst {num(r[ie]):9d}  {f}:{r[0]:>4s}  {r[1].strip()[:90]}   [{why}]")


if __name__ == "__main__":
    main()
