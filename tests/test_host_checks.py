"""CPU: the drop-in headers' host half against the reference's known-answer vectors
(tests/host_checks.cu lists them with their reference file:line), plus the delta-coded prime table
against the oracle's independent rule walk."""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

from oracle import oracle

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "_build" / "host_checks"


@pytest.fixture(scope="module")
def report():
    if not EXE.exists():
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        EXE.parent.mkdir(exist_ok=True)
        subprocess.run([nvcc, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                        "--expt-extended-lambda", "--expt-relaxed-constexpr", "-O1", "-lineinfo",
                        f"-I{ROOT / 'include'}", str(ROOT / "tests" / "host_checks.cu"), "-o", str(EXE)],
                       check=True, capture_output=True)
    res = subprocess.run([str(EXE)], capture_output=True, text=True)
    return json.loads(res.stdout)


def test_all_known_answer_checks_pass(report):
    assert report["total"] > 600
    assert report["failed"] == 0, report["failures"]


def test_prime_table_matches_rule_walk(report):
    for n, got in report["primes"].items():
        assert got == oracle.lib().oracle_prime_at_least(int(n)), n


def test_prime_table_regenerates_identically(tmp_path):
    """The committed delta table is exactly what tools/gen_prime_table.py produces from the rule."""
    out = tmp_path / "prime_table.hpp"
    subprocess.run(["python", str(ROOT / "tools" / "gen_prime_table.py"), "-o", str(out)], check=True,
                   capture_output=True)
    assert out.read_text() == (ROOT / "include" / "cuco" / "b200" / "prime_table.hpp").read_text()


def test_owner_function_matches_its_python_restatement(report):
    """exchange_owner (include/cuco/b200/bulk_kernels.cuh) is restated in Python by the gloo test's stand-in
    router (tests/test_partitioned_gloo.py::owner_of): both must agree, or the CPU multi-process tests would
    exercise a different partition than the GPU path."""
    from test_partitioned_gloo import owner_of
    assert len(report["owners"]) == 12
    for name, owner in report["owners"].items():
        key, ranks = name.split("/")
        assert owner == owner_of(int(key), 0x9E3779B97F4A7C15, int(ranks)), name
