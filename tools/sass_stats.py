#!/usr/bin/env python
"""SASS statistics of one kernel of a built library: python tools/sass_stats.py lib.so <substring of the mangled name>
Prints instruction count, the span of the SAD loop (first..last VABSDIFF4), spills inside / outside it and the opcode mix of the
main loop.  Static counts only (no GPU needed): use it before spending GPU time on a variant."""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = [ln for ln in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s", ln)]
    ops = []
    for ln in ins:
        t = ln.split("*/", 1)[1].strip().rstrip(";").split()
        if not t:
            continue
        op = t[1] if t[0].startswith("@") else t[0]
        ops.append(op)
    v = [i for i, o in enumerate(ops) if o.startswith("VABSDIFF4")]
    print(name[:90], "instructions", len(ops), "bytes", 16 * len(ops))
    if v:
        lo, hi = v[0], v[-1]
        # main loop: from the SAD pass to the backward branch after it -- approximate by [lo - 60, last BRA]
        spill_in = sum(1 for o in ops[lo:hi] if o.startswith(("LDL", "STL")))
        spill_all = sum(1 for o in ops if o.startswith(("LDL", "STL")))
        print(f"  VABSDIFF4 {len(v)} span [{lo},{hi}]  spills in SAD span {spill_in}, total {spill_all}")
        tail = collections.Counter(o.split(".")[0] for o in ops[hi:])
        import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL); print("  after SAD span:", dict(tail.most_common(14)), "total", len(ops) - hi)
        head = collections.Counter(o.split(".")[0] for o in ops[:lo])
        print("  before SAD span:", dict(head.most_common(10)), "total", lo)
