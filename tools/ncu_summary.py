"""Summarises an `ncu --page raw --csv` export: one row per profiled launch with duration, DRAM
bytes, pipe utilisation; optionally writes the dominant GEMM's per-launch DRAM traffic as JSON.

    python tools/ncu_summary.py gpurun_out/prof_gemm_trunk.raw.csv [--traffic-json profiles/r01f_gemm_traffic.json --pick 2]
"""
import argparse
import csv
import json

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--traffic-json")
    ap.add_argument("--pick", type=int, default=0, help="row whose DRAM traffic is written to --traffic-json")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    print("| # | kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    recs = []
    for i, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        recs.append(d)
        cells = []
        for k, _ in COLS:
            v = d.get(k, "")
            try:
                cells.append(f"{float(v.replace(',', '')):.1f} {u.get(k, '')}".strip())
            except ValueError:
                cells.append(v)
        print(f"| {i} | `{d['Kernel Name'][:44]}` | " + " | ".join(cells) + " |")
    if a.traffic_json:
        d = recs[a.pick]
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        json.dump({"bytes": rd + wr, "read": rd, "write": wr, "kernel": d["Kernel Name"][:60], "launch": a.note,
                   "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")), "source": a.csv}, open(a.traffic_json, "w"))


if __name__ == "__main__":
    main()
