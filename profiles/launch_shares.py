"""Per-kernel share of one step from an ncu launch list (--metrics gpu__time_duration.sum --csv).

  python profiles/launch_shares.py profiles/r01_launches_carry.csv [last_steps] [first_kernel_of_a_step]

The last `last_steps` steps are delimited by the launches of `first_kernel_of_a_step` (default k_step_begin)."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 3
first = sys.argv[3] if len(sys.argv) > 3 else "k_step_begin"
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
names = [r[4].split("(")[0] for r in rows]
ns = [float(r[14].replace(",", "")) for r in rows]
starts = [i for i, n in enumerate(names) if n.endswith(first)]
begin = starts[-last]
acc = OrderedDict()
for n, t in zip(names[begin:], ns[begin:]):
    c, s = acc.get(n, (0, 0.0))
    acc[n] = (c + 1, s + t)
tot = sum(s for _, s in acc.values())
print(f"{len(rows)} launches in the file; last {last} steps = launches {begin}..{len(rows) - 1}, {tot / last / 1e3:.1f} us/step of kernel time")
for n, (c, s) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:44s} n={c:3d} {s / last / 1e3:9.1f} us/step {100 * s / tot:5.1f}%")
