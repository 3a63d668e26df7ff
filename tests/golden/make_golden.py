"""Generates the golden fixtures in this directory by running the REFERENCE's own receiver core
(/root/reference/lib/linux_x86_64/libdabsdr.so.4.0.1, through oracle/_ref/ref_harness) over seeded synthetic IQ files.

Run in the build container (the reference is not available on the GPU box):   python tests/golden/make_golden.py
Each fixture holds: the generator parameters (so tests can rebuild the identical file), the SHA-1 of the raw file, and the
observable results of the reference: ensemble / service / component structures, SHA-1 of every delivered access unit,
the periodic statistics, and the TII notifications (ids, levels, 384-bin spectrum).
"""
import gzip
import hashlib
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

import numpy as np  # noqa: E402
import orc  # noqa: E402
from golden_cases import CASES, build_case  # noqa: E402


def main():
    assert orc.have_ref(), "reference core not built: make -C oracle ref"
    only = set(sys.argv[1:])
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        raw, res = build_case(spec)
        with tempfile.TemporaryDirectory() as td:
            path = Path(td) / f"{name}.{spec['fmt']}"
            raw.tofile(path)
            summ, ev = orc.run_ref(path, spec["fmt"], Path(td) / "log.jsonl", tii=spec.get("tii_mode"), service=spec.get("service"),
                                   extra=(["--xpad-apps"] if spec.get("xpad_apps") else []) +
                                         (["--data-scids", str(spec["data_scids"])] if spec.get("data_scids") is not None else []) +
                                         (["--spectrum-bins", str(spec["spectrum_bins"])] if spec.get("spectrum_bins") else []))
        out = {"spec": spec, "raw_sha1": hashlib.sha1(raw.tobytes()).hexdigest(), "summary": summ, "events": []}
        for e in ev:
            if e["ev"] == "audio":
                out["events"].append({"ev": "audio", "pos": e["pos"], "id": e["id"], "ascty": e["ascty"], "header": e["header"], "len": e["len"],
                                      "sha1": hashlib.sha1(bytes.fromhex(e["data"])).hexdigest()})
            elif e["ev"] == "spectrum_bins":
                out["events"].append({"ev": "spectrum_bins", "pos": e["pos"], "type": e["type"], "bins": [float("%.6g" % x) for x in e["bins"]]})
            elif e["ev"] == "tii":
                e = dict(e)
                e["spectrum"] = [round(float(x), 3) for x in e["spectrum"]]
                out["events"].append(e)
            else:
                out["events"].append(e)
        with gzip.open(HERE / f"{name}.json.gz", "wt") as f:
            json.dump(out, f)
        n_au = sum(1 for e in out["events"] if e["ev"] == "audio")
        print(name, "frames", spec["n_frames"], "AUs", n_au, "tii", [e["ids"] for e in out["events"] if e["ev"] == "tii" and e["n"]])


if __name__ == "__main__":
    main()
