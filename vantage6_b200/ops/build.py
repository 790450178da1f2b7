"""In-tree build of the native extension ``vantage6_b200/ops/_C*.so``.

Every ``.cu`` under ``csrc/`` is compiled for **sm_100a only**
(``-gencode arch=compute_100a,code=sm_100a -lineinfo``); the pybind11 module and the
symmetric-heap runtime are compiled with g++ and everything is linked into ONE shared
object that lives next to this file, so it travels with the repository snapshot to the GPU
box (a JIT cache under ``~/.cache`` would not).  nvcc cross-compiles without a GPU.

The reference has no native code at all; this extension is the
B200 data plane that replaces its REST/WebSocket transport.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
BUILD = HERE / "_build"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))
NVCC = str(CUDA_HOME / "bin" / "nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-Wno-unused-function"]


def ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def target_path() -> Path:
    return HERE / f"_C{ext_suffix()}"


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + CXX_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, log: Path | None = None) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError(f"command failed ({proc.returncode}): {' '.join(map(str, cmd))}\n{proc.stdout}")


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile (if stale) and return the path of the extension module."""
    import pybind11

    sources_cu = sorted(CSRC.glob("*.cu"))
    sources_cpp = sorted(CSRC.glob("*.cpp"))
    headers = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh"))
    out = target_path()
    stamp = BUILD / "stamp.txt"
    digest = _digest(sources_cu + sources_cpp + headers)
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return out
    BUILD.mkdir(exist_ok=True)
    py_inc = sysconfig.get_paths()["include"]
    incs = ["-I", str(CSRC), "-I", str(CUDA_HOME / "include"), "-I", pybind11.get_include(), "-I", py_inc]

    def compile_one(src: Path) -> Path:
        obj = BUILD / (src.name + ".o")
        if src.suffix == ".cu":
            cmd = [NVCC, *NVCC_FLAGS, *incs, "-c", str(src), "-o", str(obj)]
        else:
            cmd = ["g++", *CXX_FLAGS, *incs, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(f"[build] {src.name}", file=sys.stderr, flush=True)
        _run(cmd, BUILD / (src.name + ".log"))
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, sources_cu + sources_cpp))
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *map(str, objs),
            "-cudart", "static", "-Xlinker", "--exclude-libs,ALL", "-lpthread", "-ldl", "-lrt"]
    _run(link, BUILD / "link.log")
    stamp.write_text(digest)
    if verbose:
        print(f"[build] linked {out}", file=sys.stderr, flush=True)
    return out


def dump_sass(out_dir: Path | None = None) -> Path:
    """Write ``cuobjdump -sass`` of the extension (evidence for UTC*MMA / UTMALDG / multimem)."""
    out_dir = Path(out_dir or (HERE.parent.parent / "profiles"))
    out_dir.mkdir(parents=True, exist_ok=True)
    so = build(verbose=False)
    dst = out_dir / "sass_C.txt"
    with open(dst, "w") as f:
        subprocess.run([str(CUDA_HOME / "bin" / "cuobjdump"), "-sass", str(so)], stdout=f, check=True)
    return dst


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p)
