"""Build ``libqcpinn_b200.so`` (sm_100a only) in-tree with nvcc.

Usage: ``python qcpinn-convection-diffusion-qiskit_b200/build.py [--force] [--verbose]``.
The shared library is a plain C-ABI (include/qcpinn_b200.h); it does not link against torch.
"""

from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "csrc" / "build"
LIB_PATH = PKG_DIR / "libqcpinn_b200.so"
SOURCES = ["qcp_plan.cu", "qcp_point_f32.cu", "qcp_point_f64.cu", "qcp_data.cu", "qcp_state.cu",
           "qcp_reg.cu", "qcp_reg_f32.cu", "qcp_reg_f64.cu", "qcp_tile.cu", "qcp_tile_f32.cu",
           "qcp_tile_f64.cu", "qcp_plancheck.cu",
           "qcp_mlp.cu", "qcp_peer.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH_FLAGS + [
    "-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; the qcpinn_b200 CUDA library cannot be built")
    return cand


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh"))
    files.append(PKG_DIR.parent / "include" / "qcpinn_b200.h")
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


EXTRA_DEFINES: list[str] = []


def _compile(src: str, verbose: bool) -> Path:
    obj = BUILD_DIR / (Path(src).stem + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, *EXTRA_DEFINES, "-c", str(CSRC / src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        (BUILD_DIR / (Path(src).stem + ".ptxas.txt")).write_text(res.stderr)
    return obj


def build_library(force: bool = False, verbose: bool = False, out: Path | None = None,
                  defines: list[str] | None = None, only: list[str] | None = None) -> Path:
    """Compile every CUDA source for sm_100a and link the shared library (idempotent).
    ``out`` / ``defines`` build an experimental variant next to the default library; ``only``
    names the sources the defines affect (the other objects are taken from the default build)."""
    global BUILD_DIR, LIB_PATH
    main_build = BUILD_DIR
    if out is not None or defines:
        EXTRA_DEFINES[:] = [f"-D{d}" for d in (defines or [])]
        LIB_PATH = Path(out) if out is not None else LIB_PATH
        BUILD_DIR = CSRC / "build" / ("variant_" + LIB_PATH.stem)
        force = True
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    stamp = BUILD_DIR / "digest.txt"
    digest = _digest()
    record = BUILD_DIR / "build_record.json"
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        _note(record, {"reused": True, "digest": digest})
        return LIB_PATH
    import time

    t0 = time.time()
    sources = [s for s in SOURCES if (CSRC / s).exists()]
    if only:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(only))) as pool:
            fresh = dict(zip(only, pool.map(lambda s: _compile(s, verbose), only)))
        objs = [fresh.get(s, main_build / (Path(s).stem + ".o")) for s in sources]
    else:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(sources))) as pool:
            objs = list(pool.map(lambda s: _compile(s, verbose), sources))
    cmd = [_nvcc(), *ARCH_FLAGS, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    stamp.write_text(digest)
    _note(record, {"reused": False, "digest": digest, "sources": sources,
                   "seconds": round(time.time() - t0, 1)})
    return LIB_PATH


def _note(path: Path, entry: dict) -> None:
    """Append one line per build() call: which digest, compiled cold or reused."""
    import json
    import time

    entry["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
    with open(path, "a") as fh:
        fh.write(json.dumps(entry) + "\n")


def cold_compile_probe(src: str = "qcp_data.cu") -> float:
    """Compile one translation unit from scratch into a temporary object (never reused): proves on
    every ``build()`` call that nvcc, the sm_100a flags and the headers still work even when the
    digest-matched library was kept.  Returns the seconds it took."""
    import tempfile
    import time

    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        obj = Path(tmp) / "probe.o"
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not obj.exists():
            raise RuntimeError(f"cold compile probe failed on {src}:\n{res.stdout}\n{res.stderr}")
    return time.time() - t0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true", help="keep ptxas -v output in csrc/build/")
    ap.add_argument("--out", default=None, help="write an experimental variant to this path")
    ap.add_argument("-D", dest="defines", action="append", default=[], help="extra -D macro")
    ap.add_argument("--only", default=None,
                    help="comma-separated sources the defines affect (others reuse the default build)")
    ns = ap.parse_args()
    print(build_library(force=ns.force, verbose=ns.verbose, out=ns.out, defines=ns.defines,
                        only=ns.only.split(",") if ns.only else None))
    sys.exit(0)
