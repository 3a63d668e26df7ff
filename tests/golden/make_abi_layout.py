"""Records the binary layout of the reference's dabsdr.h (sizeof/offsetof/enum values) into abi_layout.json.
Run in the build container:  python tests/golden/make_abi_layout.py"""
import json
import subprocess
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = "/root/reference/lib/linux_x86_64"

with tempfile.TemporaryDirectory() as td:
    exe = Path(td) / "probe"
    subprocess.run(["gcc", f"-I{REF}", '-DABI_HEADER="dabsdr.h"', "-o", str(exe), str(HERE / "abi_probe.c")], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
layout = json.loads(out)
json.dump(layout, open(HERE / "abi_layout.json", "w"), indent=1, sort_keys=True)
print(len(layout), "entries")
