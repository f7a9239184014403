"""CPU: the two small files bench.py reads for `roofline.traffic` and for the dominant-launch share
(profiles/traffic.json, profiles/launch_shares.json) are exactly what tools/ncu_summarize.py derives
from the committed ncu launch list of the bench command - not hand-edited constants."""
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

import ncu_summarize  # noqa: E402

LAUNCHES = ROOT / "profiles" / "r02_launches_bench_native.csv"


def _medians():
    per = {}
    for launch in ncu_summarize.load(LAUNCHES):
        if "b200::" not in launch["kernel"]:
            continue
        k = ncu_summarize.short(launch["kernel"])
        per.setdefault(k, []).append((launch.get("gpu__time_duration.sum", 0.0),
                                      launch.get("dram__bytes_read.sum", 0.0) + launch.get("dram__bytes_write.sum", 0.0)))
    out = {}
    for k, rows in per.items():
        top = max(ns for ns, _ in rows)
        keep = [(ns, b) for ns, b in rows if ns >= top / 2]
        out[k] = (statistics.median(ns for ns, _ in keep), statistics.median(b for _, b in keep))
    return out


def test_traffic_and_launch_shares_come_from_the_committed_launch_list():
    medians = _medians()
    traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text())["native"]
    for kernel in ("lookup_kernel", "tile_route_kernel", "blocked_mutate_kernel"):
        assert kernel in medians, kernel
        assert traffic[kernel] == int(medians[kernel][1]), kernel
    shares = json.loads((ROOT / "profiles" / "launch_shares.json").read_text())["insert"]
    route, probe = medians["tile_route_kernel"][0], medians["blocked_mutate_kernel"][0]
    assert abs(shares["route"] - route / (route + probe)) < 1e-9
    assert abs(shares["probe"] - probe / (route + probe)) < 1e-9
    # the find pass moves about three times its algorithmic bytes (a 128-byte line per miss), the insert
    # passes stay close to theirs: the numbers DESIGN.md argues from
    n = 100_000_000
    assert 2.8 < traffic["lookup_kernel"] / (48 * n) < 3.2
    assert 1.0 < (traffic["tile_route_kernel"] + traffic["blocked_mutate_kernel"]) / (80 * n) < 1.4
