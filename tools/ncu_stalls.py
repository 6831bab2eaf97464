"""Per-instruction stall samples from an ncu report (source page, SASS): top lines + mbarrier wait summary."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address"]
for k, hi in enumerate(hdr_idx):
    h = rows[hi]
    end = hdr_idx[k + 1] if k + 1 < len(hdr_idx) else len(rows)
    si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    body = [r for r in rows[hi + 1:end] if len(r) > ie]

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    tot = sum(num(r[si]) for r in body)
    print("== kernel %d: %d samples, %d SASS lines" % (k, tot, len(body)))
    agg = {}
    for i, r in enumerate(body):
        m = re.search(r"TRYWAIT P\d, \[(R\d+)\+URZ(\+0x[0-9a-f]+)?\]", r[src])
        if m:
            key = m.group(2) or "+0x0"
            n = num(r[si]) + sum(num(body[i + d][si]) for d in (1, 2, 3, 4) if i + d < len(body))
            a = agg.setdefault(key, [0, 0])
            a[0] += n
            a[1] += num(r[ie])
    print("   mbarrier waits (offset: [samples, executions]):", agg)
    for r in sorted(body, key=lambda r: -num(r[si]))[:top]:
        print("   %s %6s %8s  %s" % (r[0][-5:], r[si], r[ie], r[src][:100]))
