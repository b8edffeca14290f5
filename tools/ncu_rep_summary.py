"""Summarise an `ncu --set full` report: one line of key metrics per profiled launch.

  python tools/ncu_rep_summary.py gpurun_out/prof.ncu-rep [--md]

Reads the report through `ncu -i ... --page raw --csv` (works without a GPU).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),                       # ns -> us (unit checked below)
    ("dram__bytes_read.sum", "dram_rd_MB", None),
    ("dram__bytes_write.sum", "dram_wr_MB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pct", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("smsp__issue_active.avg.pct", "issue_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem", 1),
    ("launch__occupancy_limit_registers", "occ_lim_regs", 1),
    ("launch__waves_per_multiprocessor", "waves", 1),
]

UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3,
              "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2:]
    return header, units, data


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    header, units, data = load(path)
    col = {h: i for i, h in enumerate(header)}
    names = [k[1] for k in KEYS]
    sep = " | " if md else "  "
    head = ["kernel", "grid", "block"] + names
    print(("| " if md else "") + sep.join(head) + (" |" if md else ""))
    if md:
        print("|" + "---|" * len(head))
    for r in data:
        kname = r[col["Kernel Name"]]
        short = kname.split("(")[0].replace("void ", "").replace("hpcs::", "")
        vals = [short[:44], r[col["Grid Size"]].replace(" ", ""), r[col["Block Size"]].replace(" ", "")]
        for key, _, scale in KEYS:
            if key not in col:
                vals.append("-")
                continue
            raw = r[col[key]].replace(",", "")
            if raw == "":
                vals.append("-")
                continue
            v = float(raw)
            u = units[col[key]]
            if u in UNIT_SCALE:
                v *= UNIT_SCALE[u]
            vals.append(f"{v:.1f}" if abs(v) < 1e5 else f"{v:.3g}")
        print(("| " if md else "") + sep.join(vals) + (" |" if md else ""))


if __name__ == "__main__":
    main()
