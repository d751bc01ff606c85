"""Wall time of ONE drop-in encode (C1 block search) -- used to A/B library builds swapped into the package directory."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import system_parity
res = system_parity.compare(["--frames", "9", "--gop", "8", "--sr", "16", "--subpel", "2", "--timeout", "300"])
print(json.dumps({"stock_s": res["stock"]["seconds"], "dropin_s": res["b200"]["seconds"], "same": res["stock"]["bitstream_md5"] == res["b200"]["bitstream_md5"]}))
