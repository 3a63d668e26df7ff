"""The drop-in boundary: include/dabsdr_abi.h must have the binary layout of the reference's lib/linux_x86_64/dabsdr.h
(recorded in tests/golden/abi_layout.json by tests/golden/make_abi_layout.py), and both shared libraries must export
every symbol their headers declare.  No GPU needed: nothing is computed."""
import ctypes as ct
import json
import re
import subprocess
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def built():
    from abracadabra_b200 import pipeline
    pipeline.build()
    return pipeline


def test_abi_layout_matches_reference_header():
    with tempfile.TemporaryDirectory() as td:
        exe = Path(td) / "probe"
        subprocess.run(["gcc", f"-I{ROOT / 'include'}", '-DABI_HEADER="dabsdr_abi.h"', "-o", str(exe), str(GOLDEN / "abi_probe.c")], check=True)
        mine = json.loads(subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout)
    ref = json.load(open(GOLDEN / "abi_layout.json"))
    assert mine == ref
    assert ref["sizeof(dabsdrNtfPeriodic_t)"] == 32          # asserted by the caller: src/radiocontrol.cpp:2407


def test_dropin_exports_the_24_entry_points(built):
    hdr = (ROOT / "include" / "dabsdr_abi.h").read_text()
    declared = sorted(set(re.findall(r"DABSDR_API\s+\w+\s+(dabsdr\w*)\s*\(", hdr)))
    assert len(declared) == 24
    assert built.SHIM_PATH.exists()
    out = subprocess.run(["nm", "-D", "--defined-only", str(built.SHIM_PATH)], check=True, capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported
    soname = subprocess.run(["readelf", "-d", str(built.SHIM_PATH)], check=True, capture_output=True, text=True).stdout
    assert "libdabsdr.so.4" in soname                        # lib/linux_x86_64/cmake/dabsdr-release.cmake
    lib = ct.CDLL(str(built.SHIM_PATH))
    v = (ct.c_uint8 * 4)()
    lib.dabsdrGetVersion(v)
    assert tuple(v[:3]) == (4, 0, 1)
    h = ct.c_void_p()
    lib.dabsdrInit.restype = ct.c_uint8
    assert lib.dabsdrInit(ct.byref(h)) == 0 and h.value      # EXIT_SUCCESS: src/radiocontrol.cpp:83
    lib.dabsdrDeinit(ct.byref(h))                            # never started: nothing to join
    assert not h.value


def test_batch_library_exports_every_declared_symbol(built):
    hdr = (ROOT / "include" / "b200dab.h").read_text()
    declared = sorted(set(re.findall(r"B200DAB_API\s+[\w\s\*]+?\b(b200dab_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = built.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert set(declared) == set(built.EXPORTS)
    assert lib.b200dab_version() >= 0x000100
