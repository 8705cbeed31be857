"""profiles/ncu_traffic.json from an `ncu --set full` report: per workload, per kernel name, the DRAM bytes of ONE
launch (dram__bytes_read.sum + dram__bytes_write.sum) - what bench.py reports as roofline.traffic.

    ncu -i X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_traffic.py raw.csv quad_concurrent profiles/r2/ncu_raw_tq_kernels.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    raw, workload, source = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                    "gpu__time_duration.sum")}
    out = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
        rd = float(r[ix["dram__bytes_read.sum"]]) * SCALE[units[ix["dram__bytes_read.sum"]]]
        wr = float(r[ix["dram__bytes_write.sum"]]) * SCALE[units[ix["dram__bytes_write.sum"]]]
        out[name] = rd + wr
        out.setdefault("detail", {})[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                                              "ncu_time_us": float(r[ix["gpu__time_duration.sum"]])}
    out["source"] = source + " (ncu --set full --clock-control none, one launch each)"
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    doc[workload] = out
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
