"""Builds cuCollections' OWN test sources against this repository's include/ tree and against the
reference's headers, unchanged, through the Catch2 stand-in under tests/catch2_shim/.

The sources stay where they are (/root/reference/tests/<suite>/*.cu, read-only, never copied);
only the executables are kept:

    tests/_build/reftests/<suite>__<name>_native      -I include
    oracle/_ref/reftests/<suite>__<name>_ref          -I /root/reference/include

Both travel to the GPU box with the snapshot (git-ignored, not gpurun-ignored), where
tests/test_reference_suites_gpu.py runs every pair and requires (a) every test case of the native
binary to pass and (b) the same PASS / SKIP lines from both. This closes SURVEY.md §8(b): "the
replacement must make the reference's own tests compile and pass by switching the include root".

    python tests/reference_suites.py [--jobs N] [--only static_map/erase_test ...]
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REFERENCE = Path("/root/reference")
SUITES = ("static_map", "static_set", "static_multiset", "utility")
NATIVE_DIR = ROOT / "tests" / "_build" / "reftests"
REF_DIR = ROOT / "oracle" / "_ref" / "reftests"
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-extended-lambda",
         "--expt-relaxed-constexpr", "-O1", "-diag-suppress", "20012,20011,177,1407,2361",
         # small executables: they are shipped to the GPU box with every snapshot
         "-Xfatbin", "-compress-all", "-cudart", "shared", "-Xlinker", "-s"]
# Sources that cannot be built here, with the reason (reported, never silently dropped).
EXCLUDED: dict[str, str] = {
    "static_map/custom_type_test": "exercises cuco::legacy::static_map (device_view API), outside SURVEY.md §8",
}
# Sources whose build against the REFERENCE headers breaks nvcc 12.9 itself ("Broken module found",
# the two-step payload spin of insert_and_find on 16-byte slots with storage<2>); the native build
# exists and is run, only the side-by-side comparison is skipped.
REFERENCE_BUILD_BROKEN = {"static_map/insert_and_find_test"}


def sources():
    out = []
    for suite in SUITES:
        for src in sorted((REFERENCE / "tests" / suite).glob("*.cu")):
            out.append((f"{suite}/{src.stem}", src))
    return out


def targets(name: str):
    flat = name.replace("/", "__")
    return NATIVE_DIR / f"{flat}_native", REF_DIR / f"{flat}_ref"


def _newest(paths):
    return max(p.stat().st_mtime for root in paths for p in ([root] if root.is_file() else root.rglob("*"))
               if p.is_file())


def _compile(job):
    name, src, include, exe = job
    exe.parent.mkdir(parents=True, exist_ok=True)
    deps = [src, include / "cuco", ROOT / "tests" / "catch2_shim", REFERENCE / "tests" / "test_utils.hpp",
            REFERENCE / "tests" / "test_utils.cuh"]
    if exe.exists() and exe.stat().st_mtime >= _newest(deps):
        return name, exe, 0.0, ""
    t0 = time.time()
    cmd = [NVCC, *FLAGS, f"-I{ROOT / 'tests' / 'catch2_shim'}", f"-I{REFERENCE / 'tests'}", f"-I{include}",
           str(src), "-o", str(exe)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if exe.exists():
            exe.unlink()
        return name, exe, time.time() - t0, res.stderr[-4000:]
    return name, exe, time.time() - t0, ""


def build(jobs: int | None = None, only=None, verbose: bool = True) -> list[str]:
    """Compiles what is stale; returns the list of failures (empty = all built)."""
    if not (REFERENCE / "tests").is_dir():
        return []  # GPU box: the prebuilt executables are used
    work = []
    for name, src in sources():
        if only and name not in only:
            continue
        if name in EXCLUDED:
            continue
        native, ref = targets(name)
        work.append((name, src, ROOT / "include", native))
        if name not in REFERENCE_BUILD_BROKEN:
            work.append((name, src, REFERENCE / "include", ref))
    failures = []
    with cf.ThreadPoolExecutor(jobs or max(1, (os.cpu_count() or 4))) as pool:
        for name, exe, secs, err in pool.map(_compile, work):
            if err:
                failures.append(f"{exe.name}: {err}")
                if verbose:
                    print(f"  FAILED {exe.name}\n{err}", file=sys.stderr)
            elif verbose and secs:
                print(f"  built {exe.name} in {secs:.0f}s", file=sys.stderr)
    return failures


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--only", nargs="*")
    args = ap.parse_args()
    bad = build(args.jobs, set(args.only) if args.only else None)
    print(f"{len(bad)} failures")
    sys.exit(1 if bad else 0)
