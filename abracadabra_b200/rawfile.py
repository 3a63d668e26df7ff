"""Raw IQ recordings as AbracaDABra reads and writes them.

* plain `.raw`: interleaved I,Q samples, unsigned 8 bit or signed 16 bit little endian (src/input/rawfileinput.h:42-46);
  the channel (hence frequency) is guessed from a `_`-separated token of the file name that is a Band III channel name
  (RawFileInput::setFile, src/input/rawfileinput.cpp:170-191).
* `.uff` / raw with XML header: a UTF-8 XML document padded with zero bytes to 2048 bytes, then the samples
  (writer: InputDeviceRecorder::startXmlHeader/finishXmlHeader, src/input/inputdevicerecorder.cpp:195-263, padding
  INPUTDEVICERECORDER_XML_PADDING = 2048, inputdevicerecorder.h:36; reader: RawFileInput::openDevice/parseXmlHeader,
  src/input/rawfileinput.cpp:86-130, :345-549).  `Datablock Count` counts I and Q values; frequency is in kHz.
"""
from __future__ import annotations

import datetime
import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import dabconst as C

XML_PADDING = 2048
_NP = {"u8": np.uint8, "s16": np.dtype("<i2")}
_CONTAINER = {"u8": ("uint8", 8), "s16": ("int16", 16)}


@dataclass
class RawInfo:
    has_xml_header: bool
    fmt: str
    sample_rate: int
    frequency_khz: int
    num_samples: int            # complex samples announced by the header (0 = unknown)
    recorder: str = ""
    device: str = ""
    time: str = ""


def frequency_from_name(path: str) -> int:
    base = os.path.basename(path).split(".")[0]          # QFileInfo::baseName(): up to the first '.'
    for tok in base.split("_"):
        if tok in C.BAND3:
            return C.BAND3[tok]
    return 0


def parse_header(head: bytes) -> Optional[RawInfo]:
    """head: the first 2048 bytes.  None if they do not start with a parsable, zero-terminated XML document."""
    z = head.find(b"\x00")
    if z < 0 or z >= XML_PADDING:
        return None
    try:
        root = ET.fromstring(head[:z].decode("utf-8"))
    except Exception:
        return None
    info = RawInfo(True, "u8", 2048000, 0, 0)
    for el in root:
        if el.tag == "Recorder":
            info.recorder = f"{el.get('Name', 'N/A')} ({el.get('Version', 'N/A')})"
        elif el.tag == "Device":
            info.device = f"{el.get('Name', 'N/A')} {el.get('Model', 'N/A')}"
        elif el.tag == "Time":
            info.time = f"{el.get('Value', 'N/A')} {el.get('Unit', 'N/A')}"
        elif el.tag == "Sample":
            for s in el:
                if s.tag == "Samplerate":
                    try:
                        sr = int(s.get("Value", "2048000"))
                    except ValueError:
                        sr = 2048000
                    info.sample_rate = sr if s.get("Unit", "Hz").lower() == "hz" else 1000 * sr
                elif s.tag == "Channels":
                    cont = s.get("Container", "uint8").lower()
                    if cont == "uint8":
                        info.fmt = "u8"
                    elif cont == "int16":
                        info.fmt = "s16"
                    else:
                        raise ValueError(f"channel container '{cont}' not supported")
                    if s.get("Ordering", "LSB") == "MSB":
                        raise ValueError("MSB ordering not supported")
        elif el.tag == "Datablocks":
            blk = el.find("Datablock")
            if blk is not None:
                try:
                    info.num_samples = int(blk.get("Count", "0")) // 2
                except ValueError:
                    info.num_samples = 0
                f = blk.find("Frequency")
                if f is not None:
                    try:
                        v = int(f.get("Value", "0"))
                    except ValueError:
                        v = 0
                    info.frequency_khz = v // 1000 if f.get("Unit", "Hz").lower() == "hz" else v
    return info


def read_raw(path: str, fmt: Optional[str] = None):
    """Returns (interleaved IQ array in the file's sample format, RawInfo)."""
    with open(path, "rb") as f:
        head = f.read(XML_PADDING)
    info = parse_header(head)
    if info is None:
        fmt = fmt or ("s16" if path.endswith(".s16") else "u8")
        info = RawInfo(False, fmt, 2048000, frequency_from_name(path), 0)
        off = 0
    else:
        off = XML_PADDING
        if not info.frequency_khz:
            info.frequency_khz = frequency_from_name(path)
    data = np.fromfile(path, dtype=_NP[info.fmt], offset=off)
    data = data[:data.size & ~1]
    if info.num_samples == 0 or 2 * info.num_samples > data.size:
        info.num_samples = data.size // 2
    return data, info


def write_uff(path: str, iq: np.ndarray, fmt: str, frequency_khz: int, device_name: str = "b200dab synth", device_model: str = "",
              recorder_version: str = "b200", when: Optional[datetime.datetime] = None) -> None:
    """Write a recording with the XML header InputDeviceRecorder produces."""
    iq = np.ascontiguousarray(iq, dtype=_NP[fmt])
    cont, bits = _CONTAINER[fmt]
    when = when or datetime.datetime.now(datetime.timezone.utc)
    root = ET.Element("SDR")
    ET.SubElement(root, "Recorder", Name="AbracaDABra", Version=recorder_version)
    ET.SubElement(root, "Device", Name=device_name, Model=device_model)
    ET.SubElement(root, "Time", Value=when.strftime("%Y-%m-%d %H:%M:%S"), Unit="UTC")
    sample = ET.SubElement(root, "Sample")
    ET.SubElement(sample, "Samplerate", Value="2048000", Unit="Hz")
    ch = ET.SubElement(sample, "Channels", Bits=str(bits), Container=cont, Ordering="LSB")
    ET.SubElement(ch, "Channel", Value="I")
    ET.SubElement(ch, "Channel", Value="Q")
    blocks = ET.SubElement(root, "Datablocks")
    blk = ET.SubElement(blocks, "Datablock", Number="1", Count=str(iq.size), Unit="Channel", Offset=str(XML_PADDING))
    ET.SubElement(blk, "Frequency", Value=str(frequency_khz), Unit="kHz")
    ET.SubElement(blk, "Modulation", Value="DAB")
    xml = b'<?xml version="1.0" encoding="utf-8"?>\n' + ET.tostring(root, encoding="utf-8")
    if len(xml) >= XML_PADDING:
        raise ValueError("XML header does not fit into 2048 bytes")
    with open(path, "wb") as f:
        f.write(xml.ljust(XML_PADDING, b"\x00"))
        iq.tofile(f)
