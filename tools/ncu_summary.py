"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list: per kernel name
launches, total / mean time and DRAM traffic.  usage: tools/ncu_summary.py file.csv [first_id last_id]"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; rows = rows[1:]
iid, iname, imet, ival, iunit = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
L = collections.OrderedDict()
for r in rows:
    k = int(r[iid])
    L.setdefault(k, {"name": r[iname]})[r[imet]] = (float(r[ival].replace(",", "")), r[iunit])
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else max(L)
def us(v):
    x, u = v
    return x * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
def mb(v):
    x, u = v
    return x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
agg = collections.OrderedDict()
tot = 0
for k, d in L.items():
    if k < lo or k > hi: continue
    nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("vst::", "")
    a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
    t = us(d["gpu__time_duration.sum"]); tot += t
    a[0] += 1; a[1] += t
    if "dram__bytes_read.sum" in d: a[2] += mb(d["dram__bytes_read.sum"]); a[3] += mb(d["dram__bytes_write.sum"])
print(f"launches {lo}..{hi}: total {tot:.1f} us")
print(f"{'kernel':44s} {'n':>4s} {'total us':>10s} {'mean us':>9s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s}")
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{nm[:44]:44s} {a[0]:4d} {a[1]:10.1f} {a[1]/a[0]:9.1f} {a[1]/tot:6.1%} {a[2]:9.1f} {a[3]:9.1f} {(a[2]+a[3])/a[1]*1e3 if a[1] else 0:7.0f}")
