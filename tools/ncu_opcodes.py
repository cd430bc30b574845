"""Executed warp instructions by SASS opcode for the (first) kernel of an ncu report with source: python tools/ncu_opcodes.py rep [rows]"""
import csv, io, subprocess, sys, collections
src = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; si = h.index("Source"); ei = h.index("Instructions Executed")
cnt = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) <= ei or not r[ei].isdigit():
        continue
    op = r[si].strip().split()
    if op and op[0].startswith("@"):
        op = op[1:]
    name = op[0].rstrip(";") if op else "?"
    base = name.split(".")[0]
    cnt[base] += int(r[ei]); tot += int(r[ei])
print(rows[0][1][:100]); print("total", tot)
nrows = float(sys.argv[2]) if len(sys.argv) > 2 else 0
for k, v in cnt.most_common(28):
    print("%-10s %12d %5.1f%%%s" % (k, v, 100.0 * v / tot, "  %6.1f per row" % (v / nrows) if nrows else ""))
