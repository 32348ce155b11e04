"""profiles/r02_traffic.json from an `ncu --set full` capture: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and
duration of the captured launch of each named kernel.  bench.py reads the file at run time for `roofline.traffic`.
    python tools/ncu_traffic.py OUT.json name=REPORT.ncu-rep:frames_per_launch [name=...]"""
import csv
import io
import json
import subprocess
import sys


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        d[h] = (v, u)
    return d


def to_bytes(vu):
    v, u = vu
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_us(vu):
    v, u = vu
    return float(v.replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]


out = {"capture": "ncu --set full --import-source on --clock-control none, one launch per kernel (cold cache, serialised)", "kernels": {}}
for spec in sys.argv[2:]:
    name, rest = spec.split("=")
    rep, frames = rest.rsplit(":", 1)
    m = metrics(rep)
    out["kernels"][name] = {
        "report": rep.split("/")[-1], "kernel": m["Kernel Name"][0], "frames_per_launch": int(frames),
        "dram_bytes_per_launch": to_bytes(m["dram__bytes_read.sum"]) + to_bytes(m["dram__bytes_write.sum"]),
        "dram_read_bytes": to_bytes(m["dram__bytes_read.sum"]), "dram_write_bytes": to_bytes(m["dram__bytes_write.sum"]),
        "duration_us": to_us(m["gpu__time_duration.sum"]),
        "tensor_pipe_active_pct": float(m.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", ("nan", ""))[0].replace(",", "") or "nan")
        if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in m else None,
    }
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
