#!/usr/bin/env python
"""SASS evidence for the hand-written kernels (runs on the CPU box: `cuobjdump -sass` of the
in-tree extension).  Writes

  profiles/sass_summary.md      per kernel: instruction count and the Blackwell-native mnemonics
                                (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA,
                                LDGMC / STG..MMR / multimem = NVLS, SYNCS = mbarrier, UTCBAR = tcgen05.commit)
  profiles/sass/<kernel>.sass   full listing of the key kernels (K1/GEMM, 2-CTA GEMM, K2, K3, K4)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KEY = re.compile(r"UTC\w*MMA\w*|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOMSWS|LDGMC|MULTIMEM|SYNCS|HMMA|REDG|ATOMG|LDG|STG|MEMBAR|CCTL|ERRBAR")
FULL = ("igemm_kernel", "gemm_bf16_kernel", "gemm2_bf16_kernel", "fedavg_round_kernel", "small_allreduce_kernel", "flash_fwd_kernel", "flash_fwd2_kernel",
        "flash_bwd_dq_kernel", "flash_bwd_dkv_kernel", "bn_stats_kernel", "bias_act_bwd_kernel", "glm_tc_kernel")


def _short_name(dem: str) -> str:
    """`void ns::kernel<(int)1, (int)2>(args...)` -> `ns::kernel<1, 2>` (template arguments kept, parameter list dropped)."""
    d = dem[5:] if dem.startswith("void ") else dem
    depth, out = 0, []
    for ch in d:
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    name = "".join(out).strip()
    return re.sub(r"\((int|bool|unsigned int)\)", "", name)


def main():
    from vantage6_b200.ops.build import build

    so = build(verbose=False)
    out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)[1:]
    os.makedirs(os.path.join(ROOT, "profiles", "sass"), exist_ok=True)
    rows = []
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        short = _short_name(dem)
        instrs = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+([^;]+);", f)
        ops = collections.Counter()
        for ins in instrs:
            op = ins.strip().lstrip("@!P0123456789 ").split()[0] if ins.strip() else ""
            m = KEY.match(op)
            if m:
                base = op.split(".")[0]
                if base in ("LDG", "STG") and ".MMR" not in op and "MC" not in op:
                    continue
                keep_full = base in ("LDGMC",) or ".MMR" in op or (base == "UTMALDG" and "IM2COL" in op)       # the im2col form of TMA is evidence by itself
                ops[op if keep_full else base] += 1
        rows.append((short, len(instrs), ops))
        for k in FULL:
            if k in short:
                tag = re.sub(r"[^A-Za-z0-9_]+", "_", short)[:80]
                with open(os.path.join(ROOT, "profiles", "sass", tag + ".sass"), "w") as fh:
                    body = "\n".join(ln[:96].rstrip() for ln in f.splitlines() if not re.match(r"^\s*/\* 0x[0-9a-f]{16} \*/\s*$", ln))
                    fh.write("Function : " + body + "\n")
    with open(os.path.join(ROOT, "profiles", "sass_summary.md"), "w") as fh:
        fh.write("# SASS summary of vantage6_b200/ops/_C*.so (sm_100a)\n\n")
        fh.write("`cuobjdump -sass`; per kernel: #instructions and counts of the mnemonics that prove the Blackwell-native path.\n\n")
        fh.write("| kernel | #instr | native mnemonics |\n|---|---|---|\n")
        for short, n, ops in sorted(rows):
            fh.write(f"| `{short[:90]}` | {n} | {', '.join(f'{k} x{v}' for k, v in sorted(ops.items()))} |\n")
    print(open(os.path.join(ROOT, "profiles", "sass_summary.md")).read()[:6000])


if __name__ == "__main__":
    main()
