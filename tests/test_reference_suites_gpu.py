"""cuCollections' own Catch2 suites (tests/static_map, static_set, static_multiset, utility of the
reference, compiled UNCHANGED through tests/catch2_shim by tests/reference_suites.py) run against this
repository's headers, next to the same sources built against the reference's headers.

Bar: every test case of the native executable passes, and both executables report the same
PASS / SKIP lines (same test-case names, same assertion counts) - i.e. switching the include root is
invisible to the reference's own tests.
"""
import subprocess
from pathlib import Path

import pytest

from reference_suites import NATIVE_DIR, REF_DIR

pytestmark = pytest.mark.gpu

NATIVE = sorted(NATIVE_DIR.glob("*_native")) if NATIVE_DIR.is_dir() else []
# large_input tests allocate tens of GB and run for minutes in the reference as well
SLOW = ("large_input",)


def _run(exe: Path, timeout: int):
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=timeout)
    verdicts = [ln for ln in res.stdout.splitlines() if ln.startswith(("PASS ", "FAIL ", "SKIP "))]
    return res.returncode, verdicts, res.stdout[-3000:] + res.stderr[-2000:]


def test_reference_suites_were_built():
    assert len(NATIVE) >= 30, "build them with `python tests/reference_suites.py` (needs /root/reference)"


@pytest.mark.parametrize("exe", NATIVE, ids=lambda p: p.name.replace("_native", ""))
def test_reference_suite(exe):
    timeout = 900 if any(s in exe.name for s in SLOW) else 300
    rc, verdicts, tail = _run(exe, timeout)
    assert rc == 0 and verdicts, tail
    assert not [v for v in verdicts if v.startswith("FAIL ")], tail
    ref = REF_DIR / exe.name.replace("_native", "_ref")
    if not ref.exists():
        pytest.skip("no build of this suite against the reference headers")
    ref_rc, ref_verdicts, ref_tail = _run(ref, timeout)
    assert ref_rc == 0, ref_tail
    assert verdicts == ref_verdicts
