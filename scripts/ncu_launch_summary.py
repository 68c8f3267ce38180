"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list by kernel.
usage: python scripts/ncu_launch_summary.py launches.csv [steps captured]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
launch = collections.OrderedDict()
for r in data:
    if len(r) > vi:
        launch.setdefault(r[0], {"name": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    m = re.match(r"(?:void )?([\w:]+)(<[^(]*>)?\(", name)
    base = m.group(1) if m else name
    tmpl = m.group(2) if (m and m.group(2)) else ""
    return (base + tmpl)[:70]


agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print("%d launches in %d step(s); per step: %.3f ms of kernel time (cold-cache, serialised: compare SHARES), DRAM %.2f GB read + %.2f GB written"
      % (len(launch), steps, tot / steps / 1e6, sum(a[2] for a in agg.values()) / steps / 1e9, sum(a[3] for a in agg.values()) / steps / 1e9))
print("%-72s %6s %9s %6s %9s %9s" % ("kernel", "n/step", "ms/step", "share", "GB read", "GB write"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6.1f %9.4f %5.1f%% %9.3f %9.3f" % (k, a[0] / steps, a[1] / steps / 1e6, 100 * a[1] / tot, a[2] / steps / 1e9, a[3] / steps / 1e9))
