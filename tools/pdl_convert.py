"""One-off source transformation (kept for the record): route every `kernel<<<grid, block, smem, stream>>>(args)` launch in
csrc/*.cu through vst::launch(...) (cudaLaunchKernelEx + programmatic stream serialisation) and open every __global__
function with vst::pdl_grid_sync() (griddepcontrol.wait + griddepcontrol.launch_dependents).

    python tools/pdl_convert.py [files...]      # rewrites in place; idempotent
Kernels listed in HAND place their own wait (after their barrier / TMEM prologue) and are skipped on the kernel side.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "video-style-transfer_b200", "csrc")
HAND = {"tapgemm_kernel", "pcgemm_kernel"}


def match_paren(s, i):
    """s[i] == '(' -> index of the matching ')'."""
    d = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            d += 1
        elif s[j] == ")":
            d -= 1
            if d == 0:
                return j
    raise ValueError("unbalanced")


def split_top(s):
    out, d, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            d += 1
        elif ch in ")]}":
            d -= 1
        if ch == "," and d == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def convert_launches(src):
    out, pos, n = "", 0, 0
    for m in re.finditer(r"<<<", src):
        i = m.start()
        if i < pos:
            continue
        # kernel expression: identifier with optional template argument list, directly before <<<
        j = i
        while j > 0 and src[j - 1].isspace():
            j -= 1
        k = j
        if src[k - 1] == ">":   # template-id: walk back to the matching '<'
            d = 0
            while True:
                k -= 1
                if src[k] == ">":
                    d += 1
                elif src[k] == "<":
                    d -= 1
                    if d == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kern = src[k:j]
        e = src.index(">>>", i)
        cfg = split_top(src[i + 3:e])
        while len(cfg) < 4:
            cfg.append("0")
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", (kern, src[a:a + 20])
        z = match_paren(src, a)
        args = src[a + 1:z].strip()
        call = "vst::launch(%s, %s" % (kern, ", ".join(cfg)) + (", " + args if args else "") + ")"
        out += src[pos:k] + call
        pos = z + 1
        n += 1
    return out + src[pos:], n


def convert_kernels(src):
    out, pos, n = "", 0, 0
    for m in re.finditer(r"__global__", src):
        i = m.end()
        # skip `void` and an optional __launch_bounds__(...)
        rest = src[i:]
        mm = re.match(r"\s*void\s*(__launch_bounds__\s*\()?", rest)
        if not mm:
            continue
        j = i + mm.end()
        if mm.group(1):
            j = match_paren(src, j - 1) + 1
        nm = re.match(r"\s*(\w+)\s*\(", src[j:])
        name = nm.group(1)
        p0 = j + nm.end() - 1
        p1 = match_paren(src, p0)
        b = p1 + 1
        while src[b].isspace():
            b += 1
        if src[b] != "{":
            continue   # a declaration
        if name in HAND or src[b + 1:b + 60].lstrip().startswith("vst::pdl_grid_sync()"):
            continue
        out += src[pos:b + 1] + "\n  vst::pdl_grid_sync();"
        pos = b + 1
        n += 1
    return out + src[pos:], n


def main():
    files = sys.argv[1:] or sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    for f in files:
        s = open(f).read()
        s, nl = convert_launches(s)
        s, nk = convert_kernels(s)
        open(f, "w").write(s)
        print("%s: %d launches, %d kernels" % (os.path.basename(f), nl, nk))


if __name__ == "__main__":
    main()
