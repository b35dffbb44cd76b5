"""Summary of a CUPTI timeline written by scripts/trace_iterations.py: span, busy time, idle gaps, time per kernel name.
    python scripts/trace_summary.py gpurun_out/trace_field.csv [top=25]"""
import collections
import csv
import re
import sys

rows = list(csv.DictReader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ev = [(r["name"], float(r["start_us"]), float(r["dur_us"])) for r in rows]
span = max(s + d for _, s, d in ev) - ev[0][1]
busy = sum(d for _, _, d in ev)
gaps = []
end = ev[0][1]
for _, s, d in ev:
    if s > end:
        gaps.append(s - end)
    end = max(end, s + d)
big = [g for g in gaps if g > 20]
print(f"{len(ev)} device activities, span {span / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {sum(gaps) / 1e3:.2f} ms "
      f"({len(big)} gaps > 20 us = {sum(big) / 1e3:.2f} ms, largest {max(gaps) if gaps else 0:.0f} us)")
by = collections.defaultdict(lambda: [0, 0.0])
for n, _, d in ev:
    k = re.sub(r"\(.*", "", n).replace("void ", "").replace("mk::", "")
    by[k][0] += 1
    by[k][1] += d
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t / 1e3:9.2f} ms {100 * t / busy:5.1f} %  {c:6d} x {t / c:8.2f} us  {k}")
