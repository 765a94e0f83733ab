"""per-source-line breakdown of an `ncu --page source --csv` dump.

  python tools/ncu_lines.py <source.csv> <object.o> <kernel-name-substring> [topn]

Joins the SASS rows of the first kernel instance (address, samples, instructions executed) with the
`//## File ... line N` annotations of `nvdisasm -g` on the cubin extracted from <object.o>, and prints the
source lines of kcf_*.cu that executed the most warp instructions / drew the most samples.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_map(obj, kernel):
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(d, cub)], text=True)
    m = {}
    cur = None
    inside = False
    stack = []
    for ln in txt.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = kernel in ln
            continue
        if not inside:
            continue
        f = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if f:
            cur = (os.path.basename(f.group(1)), int(f.group(2)))
            if f.group(3):
                cur = cur + (os.path.basename(f.group(3)), int(f.group(4)))
            continue
        a = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if a and cur:
            m[int(a.group(1), 16)] = cur
    return m


def main(path, obj, kernel, topn=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address":
            if data:
                break
            continue
        data.append(r)
    lm = line_map(obj, kernel)
    base = int(data[0][ix["Address"]], 16)

    def n(r, h):
        try:
            return int(r[ix[h]] or 0)
        except ValueError:
            return 0
    agg = {}
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        key = lm.get(off, ("?", 0))
        # attribute to the outermost kcf_screen.cu line when the instruction was inlined from a header
        k2 = key[-2:] if len(key) == 4 else key[:2]
        a = agg.setdefault(k2, [0, 0, 0])
        a[0] += n(r, "Instructions Executed")
        a[1] += n(r, "# Samples")
        a[2] += 1
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print(f"warp instructions executed {ti}, samples {ts}, SASS instructions {len(data)}")
    src = {}
    for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
        if f not in src:
            for root in (os.path.dirname(os.path.abspath(obj)), "."):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    src[f] = open(p).read().splitlines()
                    break
            else:
                src[f] = []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        print(f"{100 * a[0] / max(ti, 1):5.1f}% inst {100 * a[1] / max(ts, 1):5.1f}% smp  sass={a[2]:4d}  {f}:{l}  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
