"""Extract the DRAM traffic of the headline kernel (first lk3_fwd_tc_kernel<64,...> launch) from an ncu raw-page
export and write profiles/headline_traffic.json, which bench.py reports as roofline.traffic.
usage: python tools/ncu_headline.py profiles/r01_final_raw.csv.gz"""
import csv
import gzip
import io
import json
import os
import sys

path = sys.argv[1]
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in data:
    if "lk3_fwd_tc_kernel<64" in r[col["Kernel Name"]] or "lk3_fwd_tc_kernel" in r[col["Kernel Name"]]:
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale[units[col["dram__bytes_write.sum"]]]
        out = {"kernel": r[col["Kernel Name"]].split("(")[0], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
               "note": "one launch, ncu --set full --clock-control none; writes still resident in the 126 MB L2 at kernel "
                       "end are not counted, so the sum can be below the algorithmic bytes",
               "source": os.path.relpath(path)}
        dst = os.path.join(os.path.dirname(os.path.abspath(path)), "headline_traffic.json")
        json.dump(out, open(dst, "w"), indent=1)
        print(out)
        break
