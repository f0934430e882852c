"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the last
complete training step (delimited by the fused AdamW launches) by kernel name."""
import collections
import csv
import re
import sys

rows = [l for l in csv.reader(open(sys.argv[1], errors="ignore")) if len(l) > 10 and l[0].isdigit()]
opt = [i for i, r in enumerate(rows) if "FusedOptim" in r[4]]
# steps end with a run of optimizer launches; take the span between the last two runs
ends = [i for k, i in enumerate(opt) if k + 1 == len(opt) or opt[k + 1] != i + 1]
step = rows[ends[-2] + 1: ends[-1] + 1] if len(ends) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("at::native::", "").replace("at::", "")[:64]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"{len(step)} launches, {tot/1e3:.1f} us (serialised, cold-cache ncu timing)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:9.1f} us {v[0]:4d} {100*v[1]/tot:5.1f}%  {k}")
if len(sys.argv) > 2:
    for r in step:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("at::native::", "").replace("at::", "")[:50]
        print(r[0], name, r[8], r[-1])
