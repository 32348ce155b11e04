"""Launch-id ranges of the LAST inference forward / training step inside an ncu --csv launch list:
    python tools/profile_ranges.py infer.csv train.csv  ->  'lo_i hi_i lo_t hi_t'"""
import csv, sys


def ids(path, needle):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    iid, iname = hdr.index("ID"), hdr.index("Kernel Name")
    seen, last = [], -1
    for r in rows[1:]:
        k = int(r[iid])
        last = max(last, k)
        if needle in r[iname] and (not seen or seen[-1] != k):
            seen.append(k)
    return seen, last


pro, hi_i = ids(sys.argv[1], "prologue_x9")
adam, hi_t = ids(sys.argv[2], "adam_kernel")
lo_t = adam[-2] + 1 if len(adam) > 1 else 0
print(pro[-1], hi_i, lo_t, adam[-1])
