"""Summarise an ``ncu --metrics gpu__time_duration.sum --csv`` launch list: time and launch count per kernel."""
import collections
import csv
import re
import sys


def main(src: str, dst: str) -> None:
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "", name) if not name.startswith("void bn::") and "bn_" not in name else name
        agg[name[:90]] += v
        cnt[name[:90]] += 1
    tot = sum(agg.values())
    with open(dst, "w") as f:
        f.write(f"total {tot / 1e6:.3f} ms over {sum(cnt.values())} launches (one federated round, eager, kernels serialized under ncu)\n")
        for k, v in agg.most_common(40):
            f.write(f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}% x{cnt[k]:4d}  {k}\n")
    print(open(dst).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
