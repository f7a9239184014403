#!/usr/bin/env python3
"""Turns the ncu launch list of `bench.py` (ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --clock-control none) into the two small files bench.py reads:

  profiles/launch_shares.json   share of each launch in a multi-launch pass (insert = route + probe)
  profiles/traffic.json         measured DRAM bytes per launch of every hot-path kernel

    python tools/ncu_summarize.py <launches.csv> [native|reference] [--tag r02]

Per-launch times under ncu are cold-cache and serialised: only the SHARES are taken from them."""
import csv
import json
import re
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def load(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        try:
            value = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        launches.setdefault(int(r[idi]), {"kernel": r[ki]})[r[mi]] = value
    return [launches[i] for i in sorted(launches)]


def short(name):
    m = re.search(r"(?:b200::|detail::|cuco::)(\w+)<", name) or re.search(r"(\w+)<", name) or re.search(r"(\w+)\(", name)
    return m.group(1) if m else name[:40]


def main():
    path = sys.argv[1]
    arm = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "native"
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "r02"
    per = {}
    for l in load(path):
        if "b200::" not in l["kernel"] and "cuco::" not in l["kernel"]:
            continue  # torch's own kernels (input generation, checks) are not part of the hot path
        k = short(l["kernel"])
        per.setdefault(k, {"ns": [], "dram": []})
        per[k]["ns"].append(l.get("gpu__time_duration.sum", 0.0))
        per[k]["dram"].append(l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0))
    # the bench's kernels run on ~1e8 elements: keep launches within 2x of the kernel's longest launch
    table = {}
    for k, v in per.items():
        top = max(v["ns"])
        keep = [i for i, ns in enumerate(v["ns"]) if ns >= top / 2]
        table[k] = {"launches": len(keep), "ms_median": statistics.median(v["ns"][i] for i in keep) / 1e6,
                    "dram_bytes_median": statistics.median(v["dram"][i] for i in keep)}
    for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms_median"]):
        print(f"{k:40s} x{v['launches']:<3d} {v['ms_median']:8.3f} ms  {v['dram_bytes_median'] / 1e9:8.3f} GB")
    traffic_file = ROOT / "profiles" / "traffic.json"
    traffic = json.loads(traffic_file.read_text()) if traffic_file.exists() else {}
    traffic[arm] = {} if arm == "native" else traffic.get(arm, {})
    for k, v in table.items():
        if v["ms_median"] > 0.05:
            traffic[arm][k] = int(v["dram_bytes_median"])
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (median over the launches of "
                       "bench.py's C2 step: 100 M pairs, LF 0.5, linear_probing<1>), written by "
                       f"tools/ncu_summarize.py from profiles/{tag}_launches_bench_{arm}.csv")
    traffic_file.write_text(json.dumps(traffic, indent=1) + "\n")
    if arm == "native":
        route = next((table[k] for k in ("tile_route_kernel", "route_kernel") if k in table), None)
        probe = next((table[k] for k in ("blocked_mutate_kernel", "stream_mutate_kernel") if k in table), None)
        if route and probe:
            both = route["ms_median"] + probe["ms_median"]
            shares = {"insert": {"route": route["ms_median"] / both, "probe": probe["ms_median"] / both},
                      "source": f"profiles/{tag}_launches_bench_native.csv (ncu launch list of bench.py)"}
            (ROOT / "profiles" / "launch_shares.json").write_text(json.dumps(shares, indent=1) + "\n")
            print("insert shares:", shares["insert"])


if __name__ == "__main__":
    main()
