"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step (the last complete run of
`--per-step` launches), kernel time per family.   python scripts/launch_summary.py launches.csv [launches_per_step]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    per_step = int(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit, val = r["Metric Unit"], float(r["Metric Value"].replace(",", ""))
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        rows.append((r["Kernel Name"], us))
    if per_step:
        # the last adam kernel closes a step
        ends = [i for i, (k, _) in enumerate(rows) if "adam" in k]
        end = ends[-1] + 1
        rows = rows[end - per_step:end]
    fam = OrderedDict()
    for k, us in rows:
        name = re.sub(r"^void ", "", k)
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("b200seg::", "")
        name = name[:90]
        t = fam.setdefault(name, [0.0, 0])
        t[0] += us
        t[1] += 1
    total = sum(v[0] for v in fam.values())
    print(f"{len(rows)} launches, {total:.1f} us summed kernel time (serialised, cold-cache under ncu)")
    for name, (us, n) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
        print(f"{us:9.1f} us {100 * us / total:5.1f}%  n={n:3d}  {name}")


if __name__ == "__main__":
    main()
