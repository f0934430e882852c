"""Attribute the warp-stall samples of an ncu report (--page source --csv, SASS view)
to CUDA source lines using nvdisasm -g line info of the same cubin.

  cuobjdump -xelf all csrc/build/<file>.o; nvdisasm -g -c <cubin> > dis.txt
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_lines.py dis.txt src.csv [kernel-substring]
"""
import collections
import csv
import re
import sys

dis, src = sys.argv[1], sys.argv[2]
ins, loc = [], None
for l in open(dis):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l) and ".byte" not in l and ".dword" not in l:
        ins.append((loc, l.strip()))
rows = list(csv.reader(open(src)))
hdr, data = rows[1], rows[2:]
i_s, i_ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
i_long, i_bar = hdr.index("stall_long_sb"), hdr.index("stall_barrier")
i_short, i_wait = hdr.index("stall_short_sb"), hdr.index("stall_wait")
print(f"{len(ins)} disassembled instructions, {len(data)} profiled")
assert len(ins) >= len(data)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
for k, r in enumerate(data):
    a = agg[ins[k][0]]
    for j, i in enumerate((i_s, i_ie, i_long, i_bar, i_short, i_wait)):
        a[j] += int(r[i])
tot = sum(a[0] for a in agg.values())
toti = sum(a[1] for a in agg.values())
print(f"samples {tot}  warp instructions {toti}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{str(loc):38s} samples {a[0]:5d} {100*a[0]/tot:5.1f}%  instr {100*a[1]/toti:5.1f}%  "
          f"long_sb {a[2]:5d} barrier {a[3]:5d} short_sb {a[4]:5d} wait {a[5]:5d}")
