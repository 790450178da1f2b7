"""Registers / spills / shared memory per kernel from the `-Xptxas -v` output the in-tree build keeps under
vantage6_b200/ops/_build/*.log (run `python -m vantage6_b200.ops.build --force` first for a complete set).

    python scripts/ptxas_summary.py > profiles/ptxas_v_r2.txt
"""
import re
import subprocess
import sys
from pathlib import Path

BUILD = Path(__file__).resolve().parent.parent / "vantage6_b200" / "ops" / "_build"


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True, check=True).stdout
        return out.splitlines()
    except Exception:  # noqa: BLE001
        return list(names)


def main() -> None:
    for log in sorted(BUILD.glob("*.cu.log")):
        entries, cur = [], None
        for line in log.read_text().splitlines():
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = [m.group(1), "", ""]
                entries.append(cur)
            elif cur is not None and "bytes stack frame" in line:
                cur[1] = line.strip()
            elif cur is not None and "Used" in line:
                cur[2] = line.split(":", 1)[1].strip()
        if not entries:
            continue
        print(f"== {log.name[:-4]}")
        for name, (_, frame, used) in zip(demangle([e[0] for e in entries]), entries):
            print(f"{name[:90]:90} | {frame} {used}")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
