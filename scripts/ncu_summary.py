#!/usr/bin/env python
"""Summarise an .ncu-rep (captured on the GPU box with `ncu --set full`) into a markdown table
under profiles/.  Runs on the CPU box: `ncu -i <rep> --page raw --csv`.

    python scripts/ncu_summary.py gpurun_out/prof_gemm.ncu-rep profiles/ncu_gemm.md
"""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("smsp__cycles_active.avg", "cycles_active"),
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if out.returncode != 0:
        raise SystemExit(out.stderr[-2000:])
    rows = list(csv.reader(io.StringIO(out.stdout)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    hdr, units, rows = load(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    lines = [f"# ncu summary of `{os.path.basename(rep)}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` on one B200; read here with",
             "`ncu -i ... --page raw --csv`.  Durations under ncu are serialised/cold-cache: use them for shares and",
             "for the DRAM/tensor percentages, not as bench values.", ""]
    cols = ["kernel"] + [m[1] for m in METRICS if m[0] in idx]
    lines.append("| " + " | ".join(cols) + " |")
    lines.append("|" + "---|" * len(cols))
    for r in rows:
        name = r[idx["Kernel Name"]][:60]
        vals = [name]
        for m, short in METRICS:
            if m not in idx:
                continue
            v = r[idx[m]]
            u = units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                if short == "time":
                    v = f"{f / 1e3:.1f} us" if u in ("ns", "nsecond") else f"{f:.1f} {u}"
                elif short in ("dram_read", "dram_write", "l2_bytes"):
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                    v = f"{f * mult / 1e6:.1f} MB"
                elif short.endswith("_pct"):
                    v = f"{f:.1f}%"
                else:
                    v = f"{f:g}"
            except ValueError:
                pass
            vals.append(v)
        lines.append("| " + " | ".join(vals) + " |")
    if peaks:
        lines += ["", f"Measured peaks on this pool (MEASURED_PEAKS.json): HBM copy {peaks.get('hbm_gbs')} GB/s, "
                      f"cuBLAS bf16 {peaks.get('bf16_tflops')} TFLOP/s burst / {peaks.get('bf16_tflops_sustained')} sustained."]
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
