"""Top stall sites of one kernel from an ncu report's source page (SASS level).
    ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > src.csv ; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys


def main(path, top_n=40):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]

    def num(r, h):
        try:
            return int(float(r[idx[h]] or 0))
        except ValueError:
            return 0
    tot = sum(num(r, "# Samples") for r in data)
    print(rows[0][1][:120] if len(rows[0]) > 1 else "", "\ntotal samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]:
        s = num(r, "# Samples")
        st = sorted([(num(r, h), h[6:]) for h in stalls], reverse=True)[:2]
        print(r[idx["Address"]][-5:], f"{s:6d} {100 * s / max(tot, 1):5.1f}%  x{num(r, 'Instructions Executed'):9d}", r[idx["Source"]][:72].ljust(72), st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
