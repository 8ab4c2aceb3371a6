"""TEST INFRASTRUCTURE — reader/writer for the reference's debug dump format "ParseOggVorbis-header-v1".

The format is what ``/root/reference/src/Callbacks.cpp:146-185,318-324`` writes and what
``/root/reference/tests/compare-debug-out.py:266-366`` reads: every item is ``u32 len`` + ``len`` bytes
(native endian). The file is the raw item ``"ParseOggVorbis-header-v1"`` followed by typed items
``key, type_id (1 byte), elem_size (1 byte), data``. Entries are ``entry-name`` [``entry-channel``] ``entry-data``.

Nothing here is on the product path; the product's own dump *writer* is C++ (csrc/debug_dump.cpp).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

MAGIC = b"ParseOggVorbis-header-v1"

# Callbacks.h:55-63 (enum DataTypeId)
_DTYPES = {
    1: np.dtype("<f4"),
    2: np.dtype("<i4"),
    3: np.dtype("<u4"),
    4: np.dtype("u1"),
    5: np.dtype("u1"),  # bool, stored as one byte
    6: np.dtype("<i8"),
    7: np.dtype("<u8"),
}


class DumpReader:
    """Streaming reader over a dump held in memory."""

    def __init__(self, data: bytes):
        self.buf = memoryview(data)
        self.pos = 0
        if self._raw() != MAGIC:
            raise ValueError("not a ParseOggVorbis-header-v1 dump")
        self.decoder_name = bytes(self._typed("decoder-name")).decode()
        self.sample_rate = int(self._typed("decoder-sample-rate")[0])
        self.num_channels = int(self._typed("decoder-num-channels")[0])

    @classmethod
    def open(cls, path: str) -> "DumpReader":
        with open(path, "rb") as f:
            return cls(f.read())

    def _raw(self) -> memoryview:
        if self.pos + 4 > len(self.buf):
            raise EOFError
        (n,) = struct.unpack_from("<I", self.buf, self.pos)
        self.pos += 4
        out = self.buf[self.pos:self.pos + n]
        if len(out) != n:
            raise ValueError("truncated dump")
        self.pos += n
        return out

    def _read_typed(self) -> Tuple[str, np.ndarray]:
        key = bytes(self._raw()).decode()
        type_id = self._raw()[0]
        elem = self._raw()[0]
        raw = self._raw()
        dt = _DTYPES[type_id]
        if dt.itemsize != elem:
            raise ValueError("elem size mismatch for %s" % key)
        arr = np.frombuffer(raw, dtype=dt)
        if type_id == 5:
            arr = arr.astype(bool)
        return key, arr

    def _typed(self, expect_key: str) -> np.ndarray:
        key, arr = self._read_typed()
        if key != expect_key:
            raise ValueError("expected %r got %r" % (expect_key, key))
        return arr

    def at_eof(self) -> bool:
        return self.pos >= len(self.buf)

    def read_entry(self) -> Tuple[str, Optional[int], np.ndarray]:
        name = bytes(self._typed("entry-name")).decode()
        key, arr = self._read_typed()
        channel = None
        if key == "entry-channel":
            channel = int(arr[0])
            key, arr = self._read_typed()
        if key != "entry-data":
            raise ValueError("expected entry-data, got %r" % key)
        return name, channel, arr

    def entries(self) -> Iterator[Tuple[str, Optional[int], np.ndarray]]:
        while not self.at_eof():
            yield self.read_entry()


@dataclass
class ChannelFloor:
    floor_number: int
    ys: Optional[np.ndarray] = None            # "floor1 ys" (absent when the floor is unused)
    final_ys: Optional[np.ndarray] = None      # "floor1 final_ys"
    step2_flag: Optional[np.ndarray] = None    # "floor1 step2_flag"
    floor: Optional[np.ndarray] = None         # "floor1 floor" (len n)
    floor_outputs: Optional[np.ndarray] = None  # "floor_outputs" (len n, f32)


@dataclass
class PacketDump:
    abs_total_pos: int = 0
    expected_ending_total_pos: int = -1
    floors: Dict[int, ChannelFloor] = field(default_factory=dict)
    after_residue: Dict[int, np.ndarray] = field(default_factory=dict)
    after_envelope: Dict[int, np.ndarray] = field(default_factory=dict)
    pcm_after_mdct: Dict[int, np.ndarray] = field(default_factory=dict)
    pcm: Dict[int, np.ndarray] = field(default_factory=dict)  # emitted after finish_audio_packet

    @property
    def blocksize(self) -> int:
        return int(len(self.pcm_after_mdct[0]))


@dataclass
class StreamDump:
    decoder_name: str
    sample_rate: int
    num_channels: int
    floor_multipliers: List[int]
    floor_xs: List[np.ndarray]
    packets: List[PacketDump]

    def pcm_concat(self) -> np.ndarray:
        """[C, frames] float32: the concatenation of every "pcm" entry per channel."""
        chans = []
        for c in range(self.num_channels):
            parts = [p.pcm[c] for p in self.packets if c in p.pcm]
            chans.append(np.concatenate(parts) if parts else np.zeros(0, np.float32))
        return np.stack(chans)


def parse_dump(data: bytes) -> StreamDump:
    """Parse a whole dump of ONE stream (entry order: Callbacks.cpp push sites listed in SURVEY.md §8c)."""
    r = DumpReader(data)
    mults: List[int] = []
    xs: List[np.ndarray] = []
    # setup part
    while True:
        name, _, arr = r.read_entry()
        if name == "finish_setup":
            break
        if name == "floor1_unpack multiplier":
            mults.append(int(arr[0]))
        elif name == "floor1_unpack xs":
            xs.append(arr.astype(np.uint32))
        else:
            raise ValueError("unexpected setup entry %r" % name)
    packets: List[PacketDump] = []
    cur: Optional[PacketDump] = None
    last_finished: Optional[PacketDump] = None
    cur_floor_ch: Optional[int] = None
    for name, ch, arr in r.entries():
        if name == "start_audio_packet":
            cur = PacketDump()
            cur_floor_ch = None
        elif name == "pcm":
            # belongs to the packet that just finished (ParseOggVorbis.hpp:1270-1271, 1051)
            assert last_finished is not None
            last_finished.pcm[ch] = arr.copy()
        elif cur is None:
            raise ValueError("entry %r outside a packet" % name)
        elif name == "abs_total_pos":
            cur.abs_total_pos = int(arr[0])
        elif name == "expected_ending_total_pos":
            cur.expected_ending_total_pos = int(arr[0])
        elif name == "floor_number":
            cur.floors[ch] = ChannelFloor(floor_number=int(arr[0]))
            cur_floor_ch = ch
        elif name == "floor1 ys":
            cur.floors[cur_floor_ch].ys = arr.astype(np.uint32)
        elif name == "floor1 final_ys":
            cur.floors[cur_floor_ch].final_ys = arr.astype(np.uint32)
        elif name == "floor1 step2_flag":
            cur.floors[cur_floor_ch].step2_flag = arr.astype(bool)
        elif name == "floor1 floor":
            cur.floors[cur_floor_ch].floor = arr.astype(np.uint32)
        elif name == "floor_outputs":
            cur.floors[ch].floor_outputs = arr.copy()
        elif name == "after_residue":
            cur.after_residue[ch] = arr.copy()
        elif name == "after_envelope":
            cur.after_envelope[ch] = arr.copy()
        elif name == "pcm_after_mdct":
            cur.pcm_after_mdct[ch] = arr.copy()
        elif name == "finish_audio_packet":
            packets.append(cur)
            last_finished = cur
            cur = None
        elif name in ("floor1 fit_value unwrapped",):
            pass  # libvorbis-only hook
        else:
            raise ValueError("unknown entry %r" % name)
    return StreamDump(r.decoder_name, r.sample_rate, r.num_channels, mults, xs, packets)


def load_dump(path: str) -> StreamDump:
    with open(path, "rb") as f:
        return parse_dump(f.read())
