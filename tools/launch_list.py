"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list: our kernels, in launch order
(last N) and aggregated. Usage: python tools/launch_list.py file.csv [last_n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[0].isdigit():
        continue
    d.setdefault((int(r[0]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
ours = [(i, k, m) for (i, k), m in d.items() if "amb_" in k]
def short(k):
    return k.split("(")[0].replace("void ", "")[:44]
print("# last %d launches of our kernels (launch order)" % last_n)
for i, k, m in ours[-last_n:]:
    print("%5d %-44s %9.2f us  dram rd %9.2f MB  wr %8.2f MB" % (i, short(k), m.get("gpu__time_duration.sum", 0) / 1e3,
          m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6))
agg = collections.OrderedDict()
for i, k, m in ours:
    a = agg.setdefault(short(k), [0, 0.0])
    a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0) / 1e3
print("# all launches of our kernels: count, mean us")
for k, (n, t) in agg.items():
    print("%-44s n=%3d mean=%9.2f us" % (k, n, t / n))
