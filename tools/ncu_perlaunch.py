"""Per-launch table of an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log.
usage: tools/ncu_perlaunch.py file.csv first_id last_id"""
import csv, re, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]; rows = rows[1:]
iid, iname, imet, ival, iunit = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
L = collections.OrderedDict()
for r in rows:
    L.setdefault(int(r[iid]), {"name": r[iname]})[r[imet]] = (float(r[ival].replace(",", "")), r[iunit])
lo, hi = int(sys.argv[2]), int(sys.argv[3])
us = lambda v: v[0] * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(v[1], 1)
mb = lambda v: v[0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(v[1], 1)
print("# per launch (id, kernel, us, DRAM read MB, write MB, GB/s)")
for k, d in L.items():
    if k < lo or k > hi: continue
    nm = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("vst::", "")
    t, r, w = us(d["gpu__time_duration.sum"]), mb(d["dram__bytes_read.sum"]), mb(d["dram__bytes_write.sum"])
    print(f"{k:4d} {nm[:32]:32s} {t:8.1f} {r:9.1f} {w:9.1f} {(r + w) / t * 1e3:7.0f}")
