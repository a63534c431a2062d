"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total time, share."""
import csv, sys, collections, re
path, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
t = collections.Counter(); n = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    us = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
    name = re.sub(r"^void (fhe_b200::)?", "", r[ki]); name = re.sub(r"\((fhe_b200::)?\w+\)$", "", name)
    t[name] += us; n[name] += 1
tot = sum(t.values())
print(title); print(); print(f"launches: {sum(n.values())}, total {tot:.1f} us"); print()
print("| kernel | launches | total us | share |"); print("|---|---:|---:|---:|")
for k, v in t.most_common(14): print(f"| `{k[:90]}` | {n[k]} | {v:.1f} | {100 * v / tot:.1f}% |")
