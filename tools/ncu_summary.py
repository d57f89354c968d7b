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


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def _scaled(v, unit):
    """ncu prints durations / bytes in the unit it likes: bring them to seconds / bytes"""
    f = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return None if (v is None or f is None) else v * f


def main(path, traffic_json=None, match=None):
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
        g = lambda kname: _scaled(_num(r[col[kname]]), units[col[kname]]) if kname in col else None
        t, rd, wr = g("gpu__time_duration.sum"), g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
        if t and rd is not None and wr is not None:
            parts.append("dram=%.0f GB/s (%.3f of the measured 6555.8)" % ((rd + wr) / t / 1e9, (rd + wr) / t / 1e9 / 6555.8))
        print("  ".join(parts))
        if traffic_json and (match is None or match in name) and rd is not None:
            import json
            json.dump({"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "time_s": t,
                       "source": "ncu --set full --clock-control none, %s (one launch of the shipped variant, B = 20)" % path},
                      open(traffic_json, "w"), indent=1)
            traffic_json = None


if __name__ == "__main__":
    # python tools/ncu_summary.py x.ncu-rep [--traffic-json profiles/dominant_conv_traffic.json [kernel-name-substring]]
    tj = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    mt = sys.argv[sys.argv.index("--traffic-json") + 2] if tj and len(sys.argv) > sys.argv.index("--traffic-json") + 2 else None
    main(sys.argv[1], tj, mt)
