#!/usr/bin/env python
"""Instruction histogram of every loop (backward branch) of one kernel in libb200dab.so -- makes the per-trellis-step counts
quoted in DESIGN.md reproducible:  python tools/sass_loops.py k_viterbi"""
import collections
import re
import statistics
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "abracadabra_b200" / "lib" / "libb200dab.so"
as_json = "--json" in sys.argv
argv = [a for a in sys.argv[1:] if a != "--json"]
name = argv[0] if argv else "k_viterbi"
found = []
sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)
for blk in blocks:
    fn = blk.split("\n", 1)[0].strip()
    if name not in fn or (name == "k_viterbi" and "trace" in fn):
        continue
    ins = []
    for l in blk.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    if not as_json:
        print(f"== {fn}: {len(ins)} instructions")
    for addr, txt in ins:
        if "BRA" not in txt:
            continue
        t = re.search(r"0x([0-9a-f]+)", txt)
        if not t or int(t.group(1), 16) >= addr:
            continue
        tgt = int(t.group(1), 16)
        body = [x for a, x in ins if tgt <= a <= addr]
        c = collections.Counter()
        for x in body:
            x = re.sub(r"^@!?U?P\w+\s+", "", x)
            c[x.split()[0].split(".")[0]] += 1
        found.append((len(body), dict(c)))
        if not as_json:
            print(f"loop {tgt:#x}..{addr:#x}: {len(body)} instr", dict(c.most_common()))

if as_json:
    # k_viterbi: the innermost loops hold TWO trellis steps plus the renormalisation block (32 / 16 packed adds + shuffle + branch),
    # which runs in one of four iterations: dynamic instructions per step = (body - 3/4 renorm) / 2
    import json
    two = [n for n, c in found if c.get("VIADDMNMX") == 64 and c.get("SHFL", 0) <= 34]
    four = [n for n, c in found if c.get("VIADDMNMX") == 32 and c.get("SHFL", 0) <= 34]
    out = {"two_lane": round((statistics.mean(two) - 0.75 * 36) / 2, 1) if two else None,
           "four_lane": round((statistics.mean(four) - 0.75 * 20) / 2, 1) if four else None,
           "loops": {"two_lane_bodies": two, "four_lane_bodies": four},
           "how": "tools/sass_loops.py --json: cuobjdump -sass of libb200dab.so, innermost loops of k_viterbi (two steps + renormalisation block per body)"}
    print(json.dumps(out))
