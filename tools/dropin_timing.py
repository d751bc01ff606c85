#!/usr/bin/env python
"""Wall time of the stock encoder CLI against the same CLI linked with the drop-in (tools/system_parity.py) on a few
configurations; one JSON line each: seconds of both builds, identical streams yes / no, what went through the device."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import system_parity

CASES = {
    "C1 QCIF SR16 GOP8 SATD, block search": ["--frames", "9", "--gop", "8", "--sr", "16", "--subpel", "2"],
    "C1 QCIF SR16 GOP8 SATD, TZ search (shipped defaults)": ["--frames", "9", "--gop", "8", "--sr", "16", "--subpel", "2", "--searchmode", "4", "--elsr", "4", "--fastbi", "1"],
    "C2 QCIF->CIF SR32, block search": ["--width", "352", "--height", "288", "--layers", "2", "--frames", "5", "--gop", "4", "--sr", "32", "--subpel", "2"],
    "CIF SR64 GOP4, block search": ["--width", "352", "--height", "288", "--frames", "5", "--gop", "4", "--sr", "64", "--subpel", "2"],
    "C3-like 720p SR32 all partitions, 3 frames": ["--width", "1280", "--height", "720", "--frames", "3", "--gop", "2", "--sr", "32", "--subpel", "2", "--all-partitions"],
}

for name, args in CASES.items():
    res = system_parity.compare(args + ["--timeout", "1400"])
    same = res["stock"]["bitstream_md5"] == res["b200"]["bitstream_md5"] and res["stock"]["recon_md5"] == res["b200"]["recon_md5"]
    print(json.dumps({"case": name, "stock_s": res["stock"]["seconds"], "dropin_s": res["b200"]["seconds"], "identical": same,
                      "device_path": res["b200"]["device_path"]}), flush=True)
