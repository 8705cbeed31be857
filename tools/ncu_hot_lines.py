"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_hot_lines.py src.csv [kernel-substring] [top]
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(path)))
    i, fn, fpath = 0, None, None
    per_fn = defaultdict(lambda: defaultdict(lambda: [0, 0, defaultdict(int), ""]))
    hdr = None
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "File Path":
            fpath = r[1]
        elif r and r[0] == "Function Name":
            fn = r[1]
        elif r and r[0] == "Line No":
            hdr = r
        elif r and hdr and len(r) == len(hdr) and fn:
            d = dict(zip(hdr, r))
            try:
                samples = int(d.get("# Samples") or 0)
            except ValueError:
                samples = 0
            key = (fpath.split("/")[-1], d["Line No"])
            e = per_fn[fn][key]
            e[0] += samples
            try:
                e[1] += int(d.get("Instructions Executed") or 0)
            except ValueError:
                pass
            for k, v in d.items():
                if k.startswith("stall_") and v not in ("", "0"):
                    try:
                        e[2][k[6:]] += int(v)
                    except ValueError:
                        pass
            if not e[3]:
                e[3] = r[1][:110]
        i += 1
    for fn, lines in per_fn.items():
        if want not in fn:
            continue
        total = sum(e[0] for e in lines.values())
        print(f"== {fn[:90]}  total samples {total}")
        for key, e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            st = ", ".join(f"{k}:{v}" for k, v in sorted(e[2].items(), key=lambda kv: -kv[1])[:3])
            print(f"  {100.0 * e[0] / max(total, 1):5.1f}%  {key[0]}:{key[1]:>4}  inst {e[1]:>8}  [{st}]  {e[3].strip()}")


if __name__ == "__main__":
    main()
