"""summarise an `ncu --page source --csv` dump: stall mix and the hottest SASS instructions (first kernel instance)."""
import csv
import sys


def main(path, topn=30):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or r[0] == "Address":
            if data and (len(r) != len(hdr) or r[0] == "Address"):
                if r and r[0] in ("Kernel Name", "Address"):
                    break
            continue
        data.append(r)

    def n(r, h):
        try:
            return int(r[ix[h]] or 0)
        except ValueError:
            return 0
    tot = sum(n(r, "# Samples") for r in data)
    print("total samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(n(r, h) for r in data) for h in stalls}
    s = max(1, sum(agg.values()))
    for h, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"  {h:28s} {v:8d} {100 * v / s:5.1f}%")
    top = sorted(data, key=lambda r: -n(r, "# Samples"))[:topn]
    for r in top:
        print(f"{n(r, '# Samples'):7d} {100 * n(r, '# Samples') / max(tot, 1):5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:100]}  exec={r[ix['Instructions Executed']]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
