"""Batch band scan: one recording per channel in, the rows of AbracaDABra's scanner log out.

The reference's ScannerBackend tunes channel after channel, waits for sync, ensemble and N TII notifications, and stores one
row per detected transmitter (src/tii/scannerbackend.cpp:725-956); the CSV it saves is `;`-separated with the columns of
TxTableModel's export role (src/tii/txtablemodel.h:67-81, txtablemodel.cpp ExportRole; saveToFile scannerbackend.cpp:480-552):
    Time (UTC);Channel;Frequency [kHz];UEID;Label;Services;SNR;RF Level;Main;Sub;Level;Location;Power;Distance;Azimuth
Level is 20 log10(level / strongest level) as RadioControl computes it (src/radiocontrol.cpp:397-412).  Location / power /
distance / azimuth come from the FMLIST transmitter database, which is outside this path: left empty.
"""
from __future__ import annotations

import datetime
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import dabconst as C
from . import pipeline, rawfile, shard

HEADER = "Time (UTC);Channel;Frequency [kHz];UEID;Label;Services;SNR;RF Level;Main;Sub;Level;Location;Power;Distance;Azimuth"
_CHANNEL_OF = {v: k for k, v in C.BAND3.items()}


@dataclass
class ScanRow:
    time_utc: str
    channel: str
    frequency_khz: int
    ueid: int
    label: str
    n_services: int
    snr_db: float
    main: Optional[int]
    sub: Optional[int]
    level_db: Optional[float]

    def csv(self) -> str:
        main = "" if self.main is None else str(self.main)
        sub = "" if self.sub is None else str(self.sub)
        level = "" if self.level_db is None else f"{self.level_db:5.1f}"
        return ";".join([self.time_utc, self.channel, str(self.frequency_khz), f"{self.ueid:6X}", self.label.rstrip(), str(self.n_services),
                         f"{self.snr_db:.1f}", "", main, sub, level, "", "", "", ""])


def scan_recording(rx: pipeline.Receiver, iq: np.ndarray, fmt: str, frequency_khz: int, n_tii: int = 2,
                   when: Optional[datetime.datetime] = None) -> List[ScanRow]:
    """Decode one channel recording; rows = transmitters of the n_tii-th TII window (scanner modes 1/2/4 =
    Fast/Normal/Precise, src/tii/scannerbackend.h:142-144), or one row without ids if the ensemble carries no TII."""
    when = when or datetime.datetime.now(datetime.timezone.utc)
    dec = rx.decode(iq, fmt, tii=True, symstat=True, no_msc=True)
    if dec.n_frames == 0 or not dec.fib_ok.any():
        return []
    ens = pipeline.parse_ensemble(dec.fibs, dec.fib_ok)
    if not ens["eid"]:
        return []
    t = when.strftime("%Y-%m-%d %H:%M:%S")
    base = dict(time_utc=t, channel=_CHANNEL_OF.get(frequency_khz, ""), frequency_khz=frequency_khz, ueid=ens["ueid"], label=ens["label"],
                n_services=len(ens["services"]), snr_db=dec.snr_db)
    wins = [w for w in dec.tii if w.ids]
    if not wins:
        return [ScanRow(main=None, sub=None, level_db=None, **base)]
    ids = wins[min(n_tii, len(wins)) - 1].ids
    top = max(l for _, _, l in ids)
    return [ScanRow(main=m, sub=s, level_db=20 * math.log10(l / top), **base) for m, s, l in ids]


def scan_files(paths: Sequence[str], device: int = 0, rank: int = 0, world: int = 1, dist=None, n_tii: int = 2) -> List[str]:
    """Band scan over recordings sharded round-robin across ranks (one GPU each); returns the CSV lines on rank 0."""
    rx = pipeline.Receiver(device)

    def one(path):
        iq, info = rawfile.read_raw(path)
        return [r.csv() for r in scan_recording(rx, iq, info.fmt, info.frequency_khz, n_tii)]

    parts = shard.run_sharded(list(paths), one, rank, world, dist)
    rx.close()
    return [HEADER] + [line for part in parts for line in part]
