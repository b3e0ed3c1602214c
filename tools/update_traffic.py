#!/usr/bin/env python
"""profiles/roofline_traffic.json from an `ncu -i <rep> --page raw --csv` export of the dominant kernel.

    ncu -i gpurun_out/r2_map_vec.ncu-rep --page raw --csv > profiles/r2_map_vec_ncu_raw.csv
    python tools/update_traffic.py profiles/r2_map_vec_ncu_raw.csv [nvlink.csv ...]

Stamps the sha256 of the kernel's SOURCE (kubetorch_b200.device.lib.map_kernel_source_sha256): bench.py reports `traffic: null` with a
"stale" note when the source has changed since the capture, instead of quoting bytes of a kernel that no longer exists."""
import csv
import hashlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def source_sha():
    sys.path.insert(0, REPO)
    from kubetorch_b200.device import lib as L

    return L.map_kernel_source_sha256()


def raw_rows(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    out = []
    for r in rows[1:]:
        if r[0] in ("", "ID") or not r[0].strip('"').isdigit():
            continue
        out.append(dict(zip(hdr, r)))
    return hdr, out


def main():
    raw = sys.argv[1]
    hdr, rows = raw_rows(raw)
    rows = [r for r in rows if "map_vec_kernel" in r.get("Kernel Name", "")]
    assert rows, "no map_vec_kernel launches in " + raw

    def col(name):
        vals = [float(r[name].replace(",", "")) for r in rows if r.get(name) not in (None, "")]
        return sum(vals) / len(vals)

    unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = {}
    with open(raw) as f:
        lines = list(csv.reader(f))
    if len(lines) > 1:
        units = dict(zip(lines[0], lines[1]))
    rd = col("dram__bytes_read.sum") * unit_scale.get(units.get("dram__bytes_read.sum", "byte"), 1)
    wr = col("dram__bytes_write.sum") * unit_scale.get(units.get("dram__bytes_write.sum", "byte"), 1)
    rec = {
        "kernel": rows[0]["Kernel Name"], "workload": "x->2x over 64Mi fp32 (256 MiB in, 256 MiB out), one launch",
        "source": f"{os.path.relpath(raw, REPO)} (ncu --set full --clock-control none, {len(rows)} launches averaged)",
        "dram_bytes_read_per_launch": int(rd), "dram_bytes_write_per_launch": int(wr),
        "dram_bytes_per_launch": int(rd + wr), "algorithmic_bytes_per_launch": 536870912,
        "kernel_source_sha256": source_sha(),
    }
    for extra in sys.argv[2:]:     # NVLink counter CSVs (ncu --metrics nvl*): "key=path"
        key, path = extra.split("=", 1)
        rec[key] = json.load(open(path)) if path.endswith(".json") else path
    with open(os.path.join(REPO, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(rec, f, indent=2)
    print(json.dumps(rec, indent=2))


if __name__ == "__main__":
    main()
