"""Summarise an .ncu-rep (one `ncu --set full` capture) as a small table: per launch the duration, tensor-pipe %,
DRAM bytes, L2 / L1 / SM throughput, registers.  Run in the dev container (no GPU needed):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clk"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in reversed(list(enumerate(hdr)))}
    print("# %s  (ncu --set full --clock-control none; per launch)" % path)
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        parts = ["%-38s" % name[:38]]
        for k, short in KEYS:
            if k in col:
                v, u = r[col[k]], units[col[k]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                parts.append("%s=%s%s" % (short, v, ("" if u in ("%", "") else " " + u)))
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
