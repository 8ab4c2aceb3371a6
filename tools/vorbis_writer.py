"""A minimal Vorbis I bitstream WRITER (test infrastructure: there is no encoder in this image, and the reference ships
two fixtures only). It turns a stream description — codebooks given by their codeword lengths and VQ parameters, floor1 /
residue / mapping / mode tables — plus per-packet symbol choices (coded floor Ys, residue classifications and VQ entry
numbers) into a valid Ogg/Vorbis file, following the Vorbis I specification sections 3.2.1 (codebooks), 4.2 (headers),
4.3 (audio packets), 7.2 (floor1), 8.6 (residues) and RFC 3533 (Ogg pages).

Nothing here decodes: the files it writes are decoded by the UNMODIFIED reference (oracle/_ref/ours.bin --debug_out) to
produce golden dumps (tests/golden/make_synthetic_golden.py), and by this repo's front end + kernels under test.
No packet spans a page (the reference refuses those, src/ParseOggVorbis.hpp:89).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


def ilog(v: int) -> int:
    r = 0
    while v:
        r += 1
        v >>= 1
    return r


class BitWriter:
    """LSb-first bit packer (Vorbis I section 2)."""

    def __init__(self):
        self.acc = 0
        self.n = 0

    def put(self, value: int, bits: int):
        assert 0 <= value < (1 << bits) or bits == 0, (value, bits)
        self.acc |= value << self.n
        self.n += bits

    def put_codeword(self, code: int, length: int):
        """Huffman codewords go out most significant bit first (3.2.1)."""
        for b in range(length - 1, -1, -1):
            self.put((code >> b) & 1, 1)

    def bytes(self) -> bytes:
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def float32_pack(v: float) -> int:
    """Vorbis float32: 21-bit mantissa, 10-bit exponent biased by 788, sign (9.2.2 float32_unpack inverted)."""
    if v == 0:
        return 0
    sign = 0x80000000 if v < 0 else 0
    v = abs(v)
    exp = 0
    while v != int(v) and exp > -60:          # make the mantissa an integer
        v *= 2
        exp -= 1
    m = int(v)
    while m >= (1 << 21):
        assert m % 2 == 0, "value needs more than 21 mantissa bits"
        m //= 2
        exp += 1
    return sign | ((exp + 788) << 21) | m


def assign_codewords(lengths: Sequence[int]) -> List[Optional[int]]:
    """3.2.1: every used entry, in order, takes the lowest-valued unused codeword of its length. Linear time: avail[d] is
    the one free node at depth d (left-aligned in 32 bits) that taking earlier codewords has left over."""
    avail = [0] * 33
    codes: List[Optional[int]] = []
    first = True
    for L in lengths:
        if L == 0:
            codes.append(None)
            continue
        if first:
            first = False
            for i in range(1, L + 1):
                avail[i] = 1 << (32 - i)
            codes.append(0)
            continue
        z = L
        while z > 0 and not avail[z]:
            z -= 1
        assert z > 0, "overspecified tree"
        res = avail[z]
        avail[z] = 0
        for y in range(L, z, -1):
            avail[y] = res + (1 << (32 - y))
        codes.append(res >> (32 - L))
    return codes


def full_tree_lengths(n: int, rng: np.random.Generator, skew: float = 0.0, max_len: int = 24) -> List[int]:
    """Codeword lengths of a complete binary tree with n leaves. skew = 0: splits a random leaf (bushy); skew -> 1: keeps
    splitting the deepest leaf (long codewords, beyond any first-level decode table)."""
    assert n >= 2
    leaves = [1, 1]
    while len(leaves) < n:
        cands = [i for i, d in enumerate(leaves) if d < max_len]
        if rng.random() < skew:
            i = max(cands, key=lambda k: leaves[k])
        else:
            i = int(rng.choice(cands))
        d = leaves.pop(i)
        leaves += [d + 1, d + 1]
    out = list(leaves)
    rng.shuffle(out)
    return [int(x) for x in out]


@dataclass
class Book:
    dim: int
    lengths: List[int]                       # 0 = unused entry
    lookup_type: int = 0
    minimum: float = 0.0
    delta: float = 1.0
    value_bits: int = 4
    sequence_p: bool = False
    multiplicands: Optional[List[int]] = None
    ordered: bool = False                    # header form: lengths must be non-decreasing and all used
    codes: List[Optional[int]] = field(default_factory=list)

    def __post_init__(self):
        self.codes = assign_codewords(self.lengths)

    @property
    def n_entries(self):
        return len(self.lengths)

    def lookup1_values(self):
        r = 0
        while (r + 1) ** self.dim <= self.n_entries:
            r += 1
        return r

    def write(self, w: BitWriter):
        w.put(0x564342, 24)
        w.put(self.dim, 16)
        w.put(self.n_entries, 24)
        if self.ordered:
            assert all(l > 0 for l in self.lengths) and list(self.lengths) == sorted(self.lengths)
            w.put(1, 1)
            cur_len = self.lengths[0]
            w.put(cur_len - 1, 5)
            cur = 0
            import collections
            hist = collections.Counter(self.lengths)
            while cur < self.n_entries:
                number = hist.get(cur_len, 0)
                w.put(number, ilog(self.n_entries - cur))
                cur += number
                cur_len += 1
        else:
            w.put(0, 1)
            sparse = any(l == 0 for l in self.lengths)
            w.put(1 if sparse else 0, 1)
            for l in self.lengths:
                if sparse:
                    w.put(1 if l else 0, 1)
                    if not l:
                        continue
                w.put(l - 1, 5)
        w.put(self.lookup_type, 4)
        if self.lookup_type:
            w.put(float32_pack(self.minimum), 32)
            w.put(float32_pack(self.delta), 32)
            w.put(self.value_bits - 1, 4)
            w.put(1 if self.sequence_p else 0, 1)
            nv = self.lookup1_values() if self.lookup_type == 1 else self.n_entries * self.dim
            assert self.multiplicands is not None and len(self.multiplicands) == nv, (len(self.multiplicands or []), nv)
            for m in self.multiplicands:
                w.put(int(m), self.value_bits)

    def put(self, w: BitWriter, entry: int):
        assert self.lengths[entry] > 0, "unused entry cannot be coded"
        w.put_codeword(self.codes[entry], self.lengths[entry])

    def used_entries(self) -> List[int]:
        return [i for i, l in enumerate(self.lengths) if l]


@dataclass
class FloorClass:
    dim: int
    subclass_bits: int
    masterbook: int                           # ignored when subclass_bits == 0
    books: List[int]                          # 1 << subclass_bits entries; -1 = "no book: Y is 0"


@dataclass
class Floor1:
    partition_class: List[int]
    classes: List[FloorClass]
    multiplier: int
    rangebits: int
    xs_tail: List[int]                        # the X values after the implicit 0 and 1 << rangebits

    @property
    def xs(self):
        return [0, 1 << self.rangebits] + list(self.xs_tail)

    def write(self, w: BitWriter):
        w.put(1, 16)
        w.put(len(self.partition_class), 5)
        for c in self.partition_class:
            w.put(c, 4)
        for c in self.classes:
            w.put(c.dim - 1, 3)
            w.put(c.subclass_bits, 2)
            if c.subclass_bits:
                w.put(c.masterbook, 8)
            for b in c.books:
                w.put(b + 1, 8)
        w.put(self.multiplier - 1, 2)
        w.put(self.rangebits, 4)
        assert len(self.xs_tail) == sum(self.classes[c].dim for c in self.partition_class)
        for x in self.xs_tail:
            w.put(x, self.rangebits)


@dataclass
class Residue:
    type: int
    begin: int
    end: int
    partition_size: int
    classbook: int
    books: List[List[int]]                    # [class][pass] -> book number or -1

    @property
    def n_class(self):
        return len(self.books)

    def write(self, w: BitWriter):
        w.put(self.type, 16)
        w.put(self.begin, 24)
        w.put(self.end, 24)
        w.put(self.partition_size - 1, 24)
        w.put(self.n_class - 1, 6)
        w.put(self.classbook, 8)
        for row in self.books:
            casc = sum(1 << j for j, b in enumerate(row) if b >= 0)
            w.put(casc & 7, 3)
            if casc >> 3:
                w.put(1, 1)
                w.put(casc >> 3, 5)
            else:
                w.put(0, 1)
        for row in self.books:
            for b in row:
                if b >= 0:
                    w.put(b, 8)


@dataclass
class Mapping:
    mux: List[int]
    submap_floor: List[int]
    submap_residue: List[int]
    couplings: List[Tuple[int, int]] = field(default_factory=list)

    def write(self, w: BitWriter, channels: int):
        w.put(0, 16)
        ns = len(self.submap_floor)
        if ns > 1:
            w.put(1, 1)
            w.put(ns - 1, 4)
        else:
            w.put(0, 1)
        if self.couplings:
            w.put(1, 1)
            w.put(len(self.couplings) - 1, 8)
            bits = ilog(channels - 1)
            for m, a in self.couplings:
                w.put(m, bits)
                w.put(a, bits)
        else:
            w.put(0, 1)
        w.put(0, 2)
        if ns > 1:
            for m in self.mux:
                w.put(m, 4)
        for i in range(ns):
            w.put(0, 8)
            w.put(self.submap_floor[i], 8)
            w.put(self.submap_residue[i], 8)


@dataclass
class Mode:
    blockflag: int
    mapping: int


@dataclass
class StreamSetup:
    channels: int
    sample_rate: int
    blocksize: Tuple[int, int]
    books: List[Book]
    floors: List[Floor1]
    residues: List[Residue]
    mappings: List[Mapping]
    modes: List[Mode]

    def id_packet(self) -> bytes:
        b0, b1 = ilog(self.blocksize[0]) - 1, ilog(self.blocksize[1]) - 1
        return (b"\x01vorbis" + struct.pack("<IBIiii", 0, self.channels, self.sample_rate, 0, 0, 0) +
                bytes([b0 | (b1 << 4), 1]))

    @staticmethod
    def comment_packet() -> bytes:
        vendor = b"pov synthetic stream writer"
        return b"\x03vorbis" + struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", 0) + b"\x01"

    def setup_packet(self) -> bytes:
        w = BitWriter()
        w.put(len(self.books) - 1, 8)
        for b in self.books:
            b.write(w)
        w.put(0, 6)
        w.put(0, 16)
        w.put(len(self.floors) - 1, 6)
        for f in self.floors:
            f.write(w)
        w.put(len(self.residues) - 1, 6)
        for r in self.residues:
            r.write(w)
        w.put(len(self.mappings) - 1, 6)
        for m in self.mappings:
            m.write(w, self.channels)
        w.put(len(self.modes) - 1, 6)
        for m in self.modes:
            w.put(m.blockflag, 1)
            w.put(0, 16)
            w.put(0, 16)
            w.put(m.mapping, 8)
        w.put(1, 1)
        return b"\x05vorbis" + w.bytes()


@dataclass
class PacketChoice:
    """Everything an encoder decides for one audio packet."""
    mode: int
    prev_flag: int = 1
    next_flag: int = 1
    ys: List[Optional[List[int]]] = field(default_factory=list)      # per channel: coded Y list, or None = unused floor
    # per submap: (cls[nch][parts], entries[pass][part][ch] -> list of entry numbers) in the spec's order
    residue: List[Tuple[np.ndarray, dict]] = field(default_factory=list)
    truncate_bytes: Optional[int] = None      # cut the packet short (end-of-packet behaviour, Utils.hpp:389-392)


def propagate(used: List[bool], couplings) -> List[bool]:
    p = list(used)
    for m, a in couplings:
        if p[m] or p[a]:
            p[m] = p[a] = True
    return p


def residue_geometry(r: Residue, decode_len: int):
    lb, le = min(r.begin, decode_len), min(r.end, decode_len)
    return lb, le, (le - lb) // r.partition_size


def write_audio_packet(s: StreamSetup, pc: PacketChoice) -> bytes:
    w = BitWriter()
    w.put(0, 1)
    w.put(pc.mode, ilog(len(s.modes) - 1))
    mode = s.modes[pc.mode]
    if mode.blockflag:
        w.put(pc.prev_flag, 1)
        w.put(pc.next_flag, 1)
    mp = s.mappings[mode.mapping]
    n = s.blocksize[mode.blockflag]
    used = []
    for c in range(s.channels):
        fl = s.floors[mp.submap_floor[mp.mux[c]]]
        ys = pc.ys[c]
        used.append(ys is not None)
        if ys is None:
            w.put(0, 1)
            continue
        w.put(1, 1)
        rng_ = [256, 128, 86, 64][fl.multiplier - 1]
        yb = ilog(rng_ - 1)
        w.put(ys[0], yb)
        w.put(ys[1], yb)
        off = 2
        for cl_no in fl.partition_class:
            cl = fl.classes[cl_no]
            csub = (1 << cl.subclass_bits) - 1
            cval = 0
            if cl.subclass_bits:
                # choose, per dimension, a subclass whose book can code the Y (or "no book" for Y == 0)
                picks = []
                for i in range(cl.dim):
                    y = ys[off + i]
                    ok = [k for k, b in enumerate(cl.books) if (b < 0 and y == 0) or (b >= 0 and y < s.books[b].n_entries and s.books[b].lengths[y] > 0)]
                    assert ok, ("no subclass book can code Y", y)
                    picks.append(ok[(y + i) % len(ok)])
                for i in reversed(range(cl.dim)):
                    cval = (cval << cl.subclass_bits) | picks[i]
                s.books[cl.masterbook].put(w, cval)
            for i in range(cl.dim):
                book = cl.books[cval & csub]
                cval >>= cl.subclass_bits
                y = ys[off + i]
                if book >= 0:
                    s.books[book].put(w, y)
                else:
                    assert y == 0
            off += cl.dim
        assert off == len(ys) == len(fl.xs)
    prop = propagate(used, mp.couplings)
    for sm in range(len(mp.submap_floor)):
        chs = [c for c in range(s.channels) if mp.mux[c] == sm]
        r = s.residues[mp.submap_residue[sm]]
        cls, entries = pc.residue[sm]
        if r.type == 2:
            nch, ch_used, dl = 1, [True], len(chs) * (n // 2)
        else:
            nch, ch_used, dl = len(chs), [prop[c] for c in chs], n // 2
        lb, le, parts = residue_geometry(r, dl)
        if le == lb:
            continue
        cb = s.books[r.classbook]
        cw = cb.dim
        for pass_ in range(8):
            p = 0
            while p < parts:
                if pass_ == 0:
                    for j in range(nch):
                        if not ch_used[j]:
                            continue
                        t = 0
                        for i in range(cw):                         # first partition is the most significant digit
                            c = int(cls[j][p + i]) if p + i < parts else 0
                            t = t * r.n_class + c
                        cb.put(w, t)
                for i in range(cw):
                    if p >= parts:
                        break
                    for j in range(nch):
                        if not ch_used[j]:
                            continue
                        book = r.books[int(cls[j][p])][pass_]
                        if book < 0:
                            continue
                        for e in entries[(pass_, p, j)]:
                            s.books[book].put(w, int(e))
                    p += 1
    data = w.bytes()
    if pc.truncate_bytes is not None:
        data = data[:pc.truncate_bytes]
    return data


# ---- Ogg pages (RFC 3533) -----------------------------------------------------------------------------------
_CRC = []
for _i in range(256):
    _r = _i << 24
    for _ in range(8):
        _r = ((_r << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if _r & 0x80000000 else (_r << 1) & 0xFFFFFFFF
    _CRC.append(_r)


def ogg_crc(data: bytes) -> int:
    crc = 0
    for b in data:
        crc = ((crc << 8) & 0xFFFFFFFF) ^ _CRC[((crc >> 24) & 0xFF) ^ b]
    return crc


def ogg_page(packets: Sequence[bytes], serial: int, seq: int, granule: int, bos=False, eos=False) -> bytes:
    lacing = bytearray()
    for p in packets:
        lacing += b"\xff" * (len(p) // 255) + bytes([len(p) % 255])
    assert len(lacing) <= 255, "page would need more than 255 segments"
    hdr = bytearray(b"OggS\x00" + bytes([(2 if bos else 0) | (4 if eos else 0)]) +
                    struct.pack("<qIII", granule, serial, seq, 0) + bytes([len(lacing)]) + bytes(lacing))
    body = b"".join(packets)
    crc = ogg_crc(bytes(hdr) + body)
    hdr[22:26] = struct.pack("<I", crc)
    return bytes(hdr) + body


def write_stream(s: StreamSetup, packets: Sequence[PacketChoice], serial: int = 0x1234, packets_per_page: int = 8,
                 trim_last: int = 0) -> bytes:
    """Whole file: three header pages, then audio pages whose granule position is the number of frames decodable up to
    the page's last packet (minus trim_last on the final page: the end-of-stream trim of hpp:1028-1033)."""
    out = bytearray()
    out += ogg_page([s.id_packet()], serial, 0, 0, bos=True)
    out += ogg_page([s.comment_packet()], serial, 1, 0)
    out += ogg_page([s.setup_packet()], serial, 2, 0)
    seq = 3
    total = 0
    prev_n = 0
    datas = [write_audio_packet(s, pc) for pc in packets]
    i = 0
    while i < len(packets):
        group, gsz = [], 0
        while i < len(packets) and len(group) < packets_per_page:
            d = datas[i]
            segs = len(d) // 255 + 1
            if gsz + segs > 255:
                break
            n = s.blocksize[s.modes[packets[i].mode].blockflag]
            if prev_n:
                total += prev_n // 4 + n // 4
            prev_n = n
            group.append(d)
            gsz += segs
            i += 1
        assert group
        last = i >= len(packets)
        gran = total - (trim_last if last else 0)
        out += ogg_page(group, serial, seq, gran, eos=last)
        seq += 1
    return bytes(out)


def write_stream_spanning(s: StreamSetup, packets: Sequence[PacketChoice], serial: int = 0x1234, segments_per_page: int = 5,
                          trim_last: int = 0) -> bytes:
    """The same stream muxed the way real encoders do it (RFC 3533): the lacing values of all packets form one sequence
    that is cut into pages of `segments_per_page` segments wherever the cut falls, so packets continue across pages
    (header_type bit 0 on the continuing page). A page's granule position is the position after the last packet that ends
    on it, -1 if none does. Header packets keep their own pages (Vorbis I A.2)."""
    out = bytearray()
    seq = 0

    def emit(lacing, body, gran, cont, bos=False, eos=False):
        nonlocal seq, out
        hdr = bytearray(b"OggS\x00" + bytes([(1 if cont else 0) | (2 if bos else 0) | (4 if eos else 0)]) +
                        struct.pack("<qIII", gran, serial, seq, 0) + bytes([len(lacing)]) + bytes(lacing))
        crc = ogg_crc(bytes(hdr) + bytes(body))
        hdr[22:26] = struct.pack("<I", crc)
        out += bytes(hdr) + bytes(body)
        seq += 1

    def mux(pkts, grans, bos_first=False, eos_last=False):
        """pkts: packet byte strings; grans[i]: granule position after packet i."""
        lacing, body, cont, gran, first = [], bytearray(), False, -1, True
        for i, p in enumerate(pkts):
            segs = [255] * (len(p) // 255) + [len(p) % 255]
            at = 0
            for k, v in enumerate(segs):
                lacing.append(v)
                body += p[at:at + v]
                at += v
                if k == len(segs) - 1:
                    gran = grans[i]
                last_of_all = i == len(pkts) - 1 and k == len(segs) - 1
                if len(lacing) == segments_per_page or last_of_all:
                    emit(lacing, body, gran, cont, bos=bos_first and first, eos=eos_last and last_of_all)
                    cont = (v == 255)
                    lacing, body, gran, first = [], bytearray(), -1, False

    mux([s.id_packet()], [0], bos_first=True)
    mux([s.comment_packet()], [0])
    mux([s.setup_packet()], [0])
    datas = [write_audio_packet(s, pc) for pc in packets]
    grans, total, prev_n = [], 0, 0
    for i, pc in enumerate(packets):
        n = s.blocksize[s.modes[pc.mode].blockflag]
        if prev_n:
            total += prev_n // 4 + n // 4
        prev_n = n
        grans.append(total - (trim_last if i == len(packets) - 1 else 0))
    mux(datas, grans, eos_last=True)
    return bytes(out)


# ---------------------------------------------------------------------------------------------------------------
# page-level re-muxing of finished Ogg files (chained and multiplexed logical streams)
# ---------------------------------------------------------------------------------------------------------------
def split_pages(data: bytes) -> List[bytes]:
    """The pages of an Ogg file as byte strings."""
    pages, at = [], 0
    while at < len(data):
        assert data[at:at + 4] == b"OggS"
        nseg = data[at + 26]
        size = 27 + nseg + sum(data[at + 27:at + 27 + nseg])
        pages.append(data[at:at + size])
        at += size
    return pages


def with_serial(page: bytes, serial: int) -> bytes:
    """The same page under another stream serial number (checksum redone)."""
    p = bytearray(page)
    p[14:18] = struct.pack("<I", serial)
    p[22:26] = b"\0\0\0\0"
    p[22:26] = struct.pack("<I", ogg_crc(bytes(p)))
    return bytes(p)


def chain_files(files: Sequence[bytes], first_serial: int = 0x7000) -> bytes:
    """Logical streams one after the other (Ogg chaining), each under its own serial."""
    return b"".join(with_serial(pg, first_serial + i) for i, f in enumerate(files) for pg in split_pages(f))


def multiplex_files(files: Sequence[bytes], first_serial: int = 0x7100) -> bytes:
    """Logical streams page-interleaved (Ogg grouping): all begin-of-stream pages first, then round robin."""
    per = [[with_serial(pg, first_serial + i) for pg in split_pages(f)] for i, f in enumerate(files)]
    out = [p[0] for p in per]
    k = 1
    while any(k < len(p) for p in per):
        out += [p[k] for p in per if k < len(p)]
        k += 1
    return b"".join(out)
