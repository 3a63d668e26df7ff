#!/usr/bin/env python
"""How far do the reference core's OWN TII levels move when nothing changes but where the file starts?  Each of the first N config-3
recordings (abracadabra_b200/workloads.py: multipath, CFO, 10-30 dB, 1-8 transmitters) is decoded twice by libdabsdr 4.0.1: as is,
and with the first 2 frames + 1234 samples cut off (a different set of 20 TII nulls enters the 40-frame window, a different noise
realisation).  The level differences between the two passes bound what "level parity" with the reference can mean.

    python tools/tii_ref_spread.py [N] > profiles/rNN_tii_level_spread.json          (CPU only: needs oracle/_ref)
"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import orc
from abracadabra_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
hi, lo, rows = [], [], []
with tempfile.TemporaryDirectory() as td:
    for i in range(n):
        raw, txs, imp = W.tii_recording(i)
        got = []
        for cut in (0, 2 * 196608 + 1234):
            p = Path(td) / "t.u8"
            raw[2 * cut:].tofile(p)
            _, ev = orc.run_ref(p, "u8", Path(td) / "log.jsonl", tii=1)
            t = [e for e in ev if e["ev"] == "tii" and e["n"] > 0]
            got.append({(m, s): l for m, s, l in t[0]["ids"]} if t else {})
        a, b = got
        common = sorted(set(a) & set(b))
        top = max(a.values()) if a else 1.0
        d = {f"{m}/{s}": round(abs(20 * np.log10(a[(m, s)] / b[(m, s)])), 3) for m, s in common}
        for (m, s) in common:
            (hi if a[(m, s)] >= 0.5 * top else lo).append(abs(20 * np.log10(a[(m, s)] / b[(m, s)])))
        rows.append(dict(recording=i, snr_db=round(imp.snr_db, 1), ids_pass_1=len(a), ids_pass_2=len(b), same_id_set=set(a) == set(b), level_diff_db=d))
print(json.dumps(dict(what="reference core vs itself: same recording, start shifted by 2 frames + 1234 samples", recordings=n,
                      strong_ids=dict(n=len(hi), median_db=round(float(np.median(hi)), 3), p95_db=round(float(np.percentile(hi, 95)), 3), max_db=round(float(max(hi)), 3)),
                      weak_ids=dict(n=len(lo), median_db=round(float(np.median(lo)), 3) if lo else None, p95_db=round(float(np.percentile(lo, 95)), 3) if lo else None,
                                    max_db=round(float(max(lo)), 3) if lo else None),
                      id_sets_equal=sum(r["same_id_set"] for r in rows), rows=rows), indent=1))
