"""Per-step convolution accounting of ResNet-50 from a `conv_probe.py --groups resnet50 --time` record file.

    python scripts/conv_layer_table.py gpurun_out/conv_probe_r50.jsonl > profiles/conv_layers_r50.md

For every layer: our kernel's time (forward without / with the BatchNorm statistics epilogue, data gradient, filter
gradient), cuDNN's time for the same call on the same box, and the multiplicity; then the per-step sums -- the number a
change to the kernels has to move.  (Times are isolated launches with inputs resident in L2 where they fit; the in-round
figure is in profiles/launches_resnet50_*.)
"""
import json
import sys


def main(path: str) -> None:
    recs = [json.loads(l) for l in open(path) if l.startswith("{")]
    recs = [r for r in recs if r.get("group") == "resnet50" and "us" in r]
    rows: dict = {}
    for r in recs:
        c = r["case"]
        kind = c.split()[0]
        key = " ".join(c.split()[2:7]) if kind != "stem" else "stem"
        row = rows.setdefault(key, {"mult": r.get("mult", 1)})
        if kind == "fprop":
            tag = "fprop_stats" if "mean_err" in r else "fprop"
        else:
            tag = kind
        row[tag] = (r["us"], r["cudnn_us"])
        if kind == "stem":
            row["fprop"] = (r["us"], r["cudnn_us"])
            row["wgrad"] = (r["wgrad_us"], r["cudnn_wgrad_us"])
    tot = {k: [0.0, 0.0] for k in ("fprop", "fprop_stats", "dgrad", "wgrad")}
    print("| layer (Cin->Cout HxW k s p) | x | fprop us (cuDNN) | +stats us | dgrad us (cuDNN) | wgrad us (cuDNN) |")
    print("|---|---|---|---|---|---|")
    for key, row in rows.items():
        cells = []
        for tag in ("fprop", "fprop_stats", "dgrad", "wgrad"):
            if tag in row:
                us, cu = row[tag]
                tot[tag][0] += us * row["mult"]
                tot[tag][1] += cu * row["mult"]
                cells.append(f"{us:.1f}" if tag == "fprop_stats" else f"{us:.1f} ({cu:.1f})")
            else:
                cells.append("-")
        print(f"| {key} | {row['mult']} | " + " | ".join(cells) + " |")
    print()
    print("| per training step | ours ms | cuDNN ms | ours / cuDNN |")
    print("|---|---|---|---|")
    for tag in ("fprop", "fprop_stats", "dgrad", "wgrad"):
        o, c = tot[tag]
        if tag == "fprop_stats":
            print(f"| fprop with statistics epilogue | {o / 1e3:.3f} | (cuDNN has none; +bn_stats pass) | - |")
        elif c > 0:
            print(f"| {tag} | {o / 1e3:.3f} | {c / 1e3:.3f} | {o / c:.2f} |")
    o = tot["fprop"][0] + tot["dgrad"][0] + tot["wgrad"][0]
    c = tot["fprop"][1] + tot["dgrad"][1] + tot["wgrad"][1]
    print(f"| all three | {o / 1e3:.3f} | {c / 1e3:.3f} | {o / max(c, 1e-9):.2f} |")


if __name__ == "__main__":
    main(sys.argv[1])
