#!/usr/bin/env python
"""DRAM traffic per (column, layer) of each kernel family from an `ncu --set full` report of ONE bench step.
Usage: python profiles/extract_traffic.py report.ncu-rep COLUMNS STREAMS [out.json]
Kernels are assigned to families in launch order: k_partition_layers and the k_fast_layer_*_seg launches
that follow it form the layer family of the pass (SW first, then LW); k_fast_sweeps_* the sweep family."""
import csv
import json
import re
import subprocess
import sys

NLAY = 16


def main(path, columns, streams, out=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ik, iid = hdr.index("Kernel Name"), hdr.index("ID")
    ir, iw, it = (hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
    ur, uw = rows[1][ir], rows[1][iw]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    fam = {}
    pending = None
    for r in sorted(rows[2:], key=lambda r: int(r[iid])):
        name = r[ik]
        b = float(r[ir]) * scale[ur] + float(r[iw]) * scale[uw]
        if "k_partition_layers" in name:
            pending = b
            continue
        m = re.search(r"k_fast_(?:layer|sweeps)_(sw|lw)", name)
        if not m:
            continue
        kind = m.group(1)
        key = f"{kind}_{'sweep' if 'sweeps' in name else 'layer'}_s{streams}"
        e = fam.setdefault(key, {"dram_bytes": 0.0, "kernels": []})
        if pending is not None and "layer" in key:
            e["dram_bytes"] += pending
            e["kernels"].append("k_partition_layers")
            pending = None
        e["dram_bytes"] += b
        e["kernels"].append(name.split("(")[0])
    units = columns * NLAY
    res = {k: {"dram_bytes_per_column_layer": v["dram_bytes"] / units, "kernels": v["kernels"],
               "capture": f"{path.split('/')[-1]}: {columns} columns x {NLAY} layers, one launch per kernel"}
           for k, v in fam.items()}
    js = json.dumps(res, indent=1)
    if out:
        old = {}
        try:
            old = json.load(open(out))
        except (OSError, ValueError):
            pass
        old.update(res)
        json.dump(old, open(out, "w"), indent=1)
    print(js)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else None)
