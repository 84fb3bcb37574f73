"""Aggregate an ncu launch list (gpu__time_duration + dram bytes, --csv) by kernel name.  Usage: python tools/train_list_summary.py file.csv [--all]"""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    d = per.setdefault(r[idc], {"name": re.sub(r"^void |lass::<unnamed>::|\(.*", "", r[ki])})
    unit = r[hdr.index("Metric Unit")]
    if "time" in r[mn]: v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
    else: v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[r[mn]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
for d in per.values():
    t = d.get("gpu__time_duration.sum", 0.0); b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a = agg[d["name"]]; a[0] += 1; a[1] += t; a[2] += b; tot += t
    if "--all" in sys.argv: print("%-50s %9.1f us %8.1f MB %6.0f GB/s" % (d["name"][:50], t, b / 1e6, b / t / 1e3 if t else 0))
for n, (c, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-50s x%4d %9.1f us %9.1f MB %6.0f GB/s" % (n[:50], c, t, b / 1e6, b / t / 1e3 if t else 0))
print("total %.1f us" % tot)
