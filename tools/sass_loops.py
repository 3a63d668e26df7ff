#!/usr/bin/env python
"""Instruction histogram of every loop (backward branch) of one kernel in libb200dab.so -- makes the per-trellis-step counts
quoted in DESIGN.md reproducible:  python tools/sass_loops.py k_viterbi"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "abracadabra_b200" / "lib" / "libb200dab.so"
name = sys.argv[1] if len(sys.argv) > 1 else "k_viterbi"
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
        print(f"loop {tgt:#x}..{addr:#x}: {len(body)} instr", dict(c.most_common()))
