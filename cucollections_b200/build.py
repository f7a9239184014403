"""Builds the C-ABI shared libraries with nvcc for sm_100a (no GPU needed to build).

  libcuco_b200.so            cucollections_b200/csrc/cabi_*.cu  against  include/            (the product)
  oracle/_ref/libcuco_ref.so the same shim                      against  /root/reference/include
                             (cuco's own build: parity oracle + "cuco on the same B200" bench arm;
                              only built when the reference tree is present, i.e. in the dev
                              container - the GPU box uses the prebuilt file)

Objects are cached under cucollections_b200/_build and oracle/_build keyed by a hash of the command
line and of every header/source that can influence them, so repeated calls are cheap.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "cucollections_b200" / "csrc"
NATIVE_LIB = ROOT / "cucollections_b200" / "libcuco_b200.so"
REF_LIB = ROOT / "oracle" / "_ref" / "libcuco_ref.so"
REFERENCE_INCLUDE = Path("/root/reference/include")
NUM_KINDS = 14
# kinds whose launch parameters can be switched at run time (bench / sweep configurations)
TUNABLE_KINDS = {0, 1, 2, 3}

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
COMMON = [
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "--expt-extended-lambda",
    "--expt-relaxed-constexpr",
    "-O3",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-diag-suppress", "20012,20011,177,1407",
]


def _tree_hash(paths) -> str:
    h = hashlib.sha256()
    for root in paths:
        root = Path(root)
        files = [root] if root.is_file() else sorted(p for p in root.rglob("*") if p.is_file())
        for f in files:
            h.update(str(f).encode())
            h.update(f.read_bytes())
    return h.hexdigest()


def _compile(job):
    src, obj, flags, stamp = job
    stamp_file = obj.with_suffix(".stamp")
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj, 0.0, ""
    cmd = [NVCC, *COMMON, *flags, "-c", str(src), "-o", str(obj)]
    import time
    t0 = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    stamp_file.write_text(stamp)
    return obj, time.time() - t0, res.stderr


def _build_lib(lib: Path, build_dir: Path, include: Path, extra: list[str], reference: bool, verbose: bool,
               only_kinds=None):
    build_dir.mkdir(parents=True, exist_ok=True)
    lib.parent.mkdir(parents=True, exist_ok=True)
    # each object depends on its own source, the shared shim header and the header tree it
    # instantiates; cabi_core.cu additionally on the C header it implements
    headers = _tree_hash([include / "cuco", CSRC / "cabi_table.hpp"])
    core_hash = headers + _tree_hash([CSRC / "cabi_core.cu", ROOT / "include" / "cuco_b200.h",
                                      ROOT / "include" / "cuco" / "b200"])
    kind_hash = headers + _tree_hash([CSRC / "cabi_kind.cu"])
    jobs = []
    base = [f"-I{include}", *extra]
    jobs.append((CSRC / "cabi_core.cu", build_dir / "cabi_core.o", base, core_hash))
    for k in range(NUM_KINDS):
        flags = [*base, f"-DCUCO_SHIM_KIND={k}"]
        if only_kinds is not None and k not in only_kinds:
            flags.append("-DCUCO_SHIM_STUB=1")
        elif not reference and k in TUNABLE_KINDS:
            flags.append("-DCUCO_B200_TUNABLE=1")
        jobs.append((CSRC / "cabi_kind.cu", build_dir / f"cabi_kind_{k}.o", flags, kind_hash))
    jobs = [(s, o, f, hashlib.sha256((h + " ".join(map(str, f)) + " ".join(COMMON)).encode()).hexdigest())
            for (s, o, f, h) in jobs]
    workers = max(1, min(len(jobs), (os.cpu_count() or 4)))
    objs = []
    with cf.ThreadPoolExecutor(workers) as pool:
        for obj, secs, err in pool.map(_compile, jobs):
            objs.append(obj)
            if verbose and secs:
                print(f"  built {obj.name} in {secs:.1f}s", file=sys.stderr)
    newest = max(o.stat().st_mtime for o in objs)
    if not lib.exists() or lib.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-o", str(lib), *map(str, objs), "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed: {' '.join(cmd)}\n{res.stderr}")
        if verbose:
            print(f"  linked {lib}", file=sys.stderr)
    return lib


def build_native(verbose: bool = False) -> Path:
    """Compiles the product library (hand-written sm_100a kernels behind the cuco:: surface)."""
    return _build_lib(NATIVE_LIB, ROOT / "cucollections_b200" / "_build", ROOT / "include", [], False, verbose)


def build_dev(kinds, verbose: bool = True) -> Path:
    """Development build: only `kinds` are real, the others are stubs; written next to the product
    library as libcuco_b200_dev.so (select it with CUCO_B200_LIB=<path>)."""
    return _build_lib(ROOT / "cucollections_b200" / "libcuco_b200_dev.so",
                      ROOT / "cucollections_b200" / "_build_dev", ROOT / "include", [], False, verbose,
                      only_kinds=set(kinds))


def build_reference(verbose: bool = False) -> Path | None:
    """Compiles cuco's own headers behind the same shim into oracle/_ref (dev container only)."""
    if not REFERENCE_INCLUDE.is_dir():
        return REF_LIB if REF_LIB.exists() else None
    return _build_lib(REF_LIB, ROOT / "oracle" / "_build", REFERENCE_INCLUDE,
                      ["-DCUCO_SHIM_REFERENCE=1"], True, verbose)


if __name__ == "__main__":
    which = sys.argv[1:] or ["native", "reference"]
    if which and which[0] == "dev":
        print(build_dev([int(k) for k in which[1:]] or [1]))
        sys.exit(0)
    if "native" in which:
        print(build_native(verbose=True))
    if "reference" in which:
        print(build_reference(verbose=True))
