"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (count, total, share).
usage: python tools/launch_summary.py launches.csv [first_id last_id] > profiles/<name>.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path) as f:
    for r in csv.reader(l for l in f if l.startswith('"')):
        if r[0] == "ID":
            continue
        rows.append((int(r[0]), r[4], r[7], r[8], float(r[14])))
rows = [r for r in rows if lo <= r[0] <= hi]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) < 110 else name[:107] + "..."


tot = defaultdict(lambda: [0, 0.0, 0.0])
for _, k, _, _, ns in rows:
    t = tot[short(k)]
    t[0] += 1
    t[1] += ns
    t[2] = max(t[2], ns)
total = sum(v[1] for v in tot.values())
ours = sum(v[1] for k, v in tot.items() if "lla_" in k)
print(f"launches {len(rows)} (ids {rows[0][0]}..{rows[-1][0]}), summed device time {total / 1e6:.2f} ms; llacuda kernels {ours / total:.1%} of it\n")
print("| kernel | launches | total ms | share | longest us |")
print("|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / total:.1%} | {v[2] / 1e3:.1f} |")
