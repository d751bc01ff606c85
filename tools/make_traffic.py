#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch of K1 / K2 / K3 from the summaries of tools/summarize_profile.py.
    python tools/make_traffic.py <pairs> ncu_k1.json ncu_k2.json ncu_k3.json > traffic.json
bench.py reports these as roofline.traffic when it runs the batch size that was profiled."""
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def to_bytes(s):
    v, u = s.split()
    return float(v) * UNIT[u]


pairs = int(sys.argv[1])
names = ["k1_halfpel", "k2_full_search", "k3_subpel"]
out = {"pairs": pairs, "dram_bytes_per_launch": {}, "kernels": {},
       "note": "one `ncu --set full --clock-control none` capture per kernel of `bench.py --pairs %d`; small launches are absorbed by the 126 MB L2 "
               "(write-backs after the kernel ends are not in the kernel's counters)" % pairs}
for name, path in zip(names, sys.argv[2:]):
    k = json.load(open(path))["kernels"][0]
    out["dram_bytes_per_launch"][name] = int(to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"]))
    out["kernels"][name] = k["kernel"].split("(")[0]
print(json.dumps(out, indent=1))
