"""profiles/ncu_traffic.json from an `ncu --page raw --csv` export of the --set full capture of THIS build:
python scripts/make_ncu_traffic.py profiles/<tag>_ncu_full_raw.csv"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = {}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ki])
        name = re.sub(r"^(\w+::)*", "", name)
        name = re.match(r"\w+", name).group(0)
        b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
        out.setdefault(name, []).append(b)
    d = {"source_hash": bench.source_hash(),
         "capture": "%s: ncu --set full --clock-control none, scripts/run_fused_once.py (batch 4, P=24000, N=200), "
                    "dram__bytes_read.sum + dram__bytes_write.sum per launch" % os.path.relpath(path, ROOT),
         "kernels": {k: int(sum(v) / len(v)) for k, v in out.items()}}
    json.dump(d, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
