"""Host front-end (Ogg framing, header/setup parse, Huffman + residue walk -> descriptors) against the reference's
dumps, on the CPU: the descriptors are applied by the pinned oracle, so what is checked here is exactly what the
host emits — coded floor Y lists, classifications and VQ entry numbers (bit-exact `after_residue`), window flags,
emit counts and granule trimming."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from parseoggvorbis_b200 import abi, lib
from tests import oracle_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = {"stereo44khz": "test.stereo44khz.ogg", "mono44khz": "test.mono44khz.ogg"}
from tests.conftest import SYNTHETIC  # noqa: E402
FIX.update({n: n + ".ogg" for n in SYNTHETIC})


def _load(name):
    with open(os.path.join(ROOT, "tests", "golden", FIX[name]), "rb") as f:
        return f.read()


@pytest.mark.parametrize("name", list(FIX))
def test_descriptors_reproduce_reference_dump(golden, name):
    g = golden[name]
    po = lib.ParsedOgg(_load(name))
    assert po.n_streams == 1
    s, b = po.get(0)
    Cn = int(g["channels"])
    assert s.channels == Cn and s.sample_rate == int(g["sample_rate"])
    assert b.n_packets == len(g["blocksize"])
    pk = np.ctypeslib.as_array(C.cast(b.packets, C.POINTER(C.c_uint8)), shape=(b.n_packets * C.sizeof(abi.pov_packet),))
    pk = pk.view(abi.PACKET_DTYPE)
    assert np.array_equal(pk["emit_frames"], g["emit_frames"])
    used = (pk["floor_used"][:, None] >> np.arange(Cn)[None, :]) & 1
    assert np.array_equal(used.astype(bool), g["floor_used"])
    # coded Y lists, bit-exact ("floor1 ys", the one floor field the reference's own harness compares)
    ys = np.ctypeslib.as_array(b.ys, shape=(b.n_ys,))
    for p in range(b.n_packets):
        yo = int(pk["ys_off"][p])
        for c in range(Cn):
            if g["floor_used"][p, c]:
                k = int(g["floor_nposts"][g["floor_number"][p, c]])
                assert np.array_equal(ys[yo:yo + k], g["ys"][p, c, :k]), (p, c)
                yo += k
    # everything downstream through the oracle
    pcm, status, cap = ob.synth_batch_raw(s, b, imdct="reference" if ob.reference_lib() is not None else "fast", capture=True)
    assert not status.any()
    oh = of = 0
    for p, n in enumerate(g["blocksize"]):
        n = int(n)
        for c in range(Cn):
            assert np.array_equal(cap["after_residue"][p, c, :n // 2], g["after_residue"][oh:oh + n // 2]), (p, c)
            assert np.array_equal(cap["after_envelope"][p, c, :n // 2], g["after_envelope"][oh:oh + n // 2]), (p, c)
            oh += n // 2
            of += n
    out = pcm.reshape(Cn, -1)
    assert out.shape == g["pcm"].shape
    assert np.abs(out - g["pcm"]).max() <= 1e-5
    if ob.reference_lib() is not None and name != "synth_shared_submap":      # (that one's golden PCM is libvorbis', another IMDCT)
        assert np.array_equal(out, g["pcm"])
    po.close()


def _expect_error(data, needle=None):
    with pytest.raises(lib.PovError) as ei:
        lib.ParsedOgg(bytes(data))
    assert ei.value.code == abi.POV_ERR_STREAM
    if needle:
        assert needle in ei.value.msg, ei.value.msg


def test_malformed_streams_fail_like_the_reference():
    data = bytearray(_load("mono44khz"))
    bad = bytearray(data); bad[0] = ord("X")
    _expect_error(bad, "capture pattern")                       # hpp:77
    bad = bytearray(data); bad[len(bad) // 2] ^= 0x55
    _expect_error(bad, "CRC")                                   # hpp:98
    _expect_error(data[:len(data) - 100], "truncated")          # hpp:90
    assert lib.ParsedOgg(bytes(data[:20])).n_streams == 0       # < 27 bytes: plain EOF (hpp:69-74)
    assert lib.ParsedOgg(b"").n_streams == 0


@pytest.mark.skipif(ob.reference_lib() is None, reason="oracle/_ref not built")
def test_error_behaviour_matches_reference_library():
    """Same inputs through the reference's own C API (hpp:1493): both accept or both reject."""
    ref = ob.reference_lib()
    ref.ogg_vorbis_full_read_from_memory.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p)]
    rng = np.random.default_rng(0)
    base = _load("mono44khz")
    cases = [base, base[:5000], base[:27], base[:4000] + base[4100:]]
    for _ in range(6):
        b = bytearray(base)
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    for data in cases:
        err = C.c_char_p(None)
        ref_rc = ref.ogg_vorbis_full_read_from_memory(data, len(data), C.byref(err))
        try:
            lib.ParsedOgg(data).close()
            ours_rc = 0
        except lib.PovError:
            ours_rc = 1
        assert (ref_rc != 0) == (ours_rc != 0), (len(data), err.value)


def _ogg_crc(page: bytes) -> int:
    crc = 0
    for byte in page:     # poly 0x04c11db7, MSB first, init 0 (hpp:92-98)
        crc ^= byte << 24
        for _ in range(8):
            crc = ((crc << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if crc & 0x80000000 else (crc << 1) & 0xFFFFFFFF
    return crc


def test_more_channels_than_the_descriptors_hold_are_refused_in_the_id_header():
    """The reference takes any uint8 channel count (hpp:107); this build's descriptors hold 8. A stream that announces
    more must be refused at its id header, before an audio packet is walked with per-channel arrays of 8."""
    data = bytearray(_load("mono44khz"))
    nseg = data[26]
    body = sum(data[27:27 + nseg])
    first = 27 + nseg                     # id packet: 0x01 "vorbis" version(4) channels(1) ...
    assert data[first:first + 7] == b"\x01vorbis"
    data[first + 11] = 9
    page = bytearray(data[:first + body]); page[22:26] = b"\0\0\0\0"
    data[22:26] = _ogg_crc(bytes(page)).to_bytes(4, "little")
    _expect_error(data, "9 channels")


def test_crc_known_answer():
    # Ogg CRC of the first page of the bundled fixture equals the value stored in its header (hpp:92-98)
    data = _load("stereo44khz")
    nseg = data[26]
    body = sum(data[27:27 + nseg])
    page = bytearray(data[:27 + nseg + body])
    stored = int.from_bytes(page[22:26], "little")
    page[22:26] = b"\0\0\0\0"
    crc = 0
    for byte in page:     # bitwise restatement: poly 0x04c11db7, MSB first, init 0
        crc ^= byte << 24
        for _ in range(8):
            crc = ((crc << 1) ^ 0x04C11DB7) & 0xFFFFFFFF if crc & 0x80000000 else (crc << 1) & 0xFFFFFFFF
    assert crc == stored


def test_scalar_book_used_as_vq_book_is_refused_like_the_reference():
    """synth_bad_vq_book.ogg: a residue class points at a lookup-type-0 codebook; the reference's decodeVector CHECK fails
    (hpp:369-370, 748 — verified when the fixture was generated). The host walk must refuse the file the same way."""
    with open(os.path.join(ROOT, "tests", "golden", "synth_bad_vq_book.ogg"), "rb") as f:
        data = f.read()
    _expect_error(data, "invalid VQ entry")


def _spanning_twin(name):
    """The synthetic stream `name` muxed with packets continuing across pages (tools/vorbis_writer.write_stream_spanning)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_synthetic_golden as msg
    import vorbis_writer as vw
    fn, seed, _ = msg.SCENARIOS[name]
    s, packets, trim = fn(np.random.default_rng(seed))
    plain = vw.write_stream(s, packets, serial=0x5000 + seed, packets_per_page=6, trim_last=trim)
    with open(os.path.join(ROOT, "tests", "golden", "synth_%s.ogg" % name), "rb") as f:
        assert f.read() == plain, "the committed fixture is what the generator writes"
    return vw.write_stream_spanning(s, packets, serial=0x5000 + seed, segments_per_page=5, trim_last=trim)


@pytest.mark.parametrize("name", ["two_submaps", "codebooks"])
def test_packets_spanning_pages(golden, name):
    """f4 (hpp:89): refused by default, like the reference; with POV_PARSE_ALLOW_SPANNING the same packets muxed across page
    boundaries decode to exactly what the reference decodes from the non-spanning file."""
    data = _spanning_twin(name)
    _expect_error(data, "spanning pages")
    g = golden["synth_" + name]
    po = lib.ParsedOgg(data, allow_spanning=True)
    s, b = po.get(0)
    assert b.n_packets == len(g["blocksize"])
    pcm, status, cap = ob.synth_batch_raw(s, b, imdct="reference" if ob.reference_lib() is not None else "fast", capture=True)
    assert not status.any()
    out = pcm.reshape(int(g["channels"]), -1)
    assert out.shape == g["pcm"].shape and np.abs(out - g["pcm"]).max() <= 1e-5
    oh = 0
    for p, n in enumerate(g["blocksize"]):
        for c in range(int(g["channels"])):
            assert np.array_equal(cap["after_residue"][p, c, :int(n) // 2], g["after_residue"][oh:oh + int(n) // 2]), (p, c)
            oh += int(n) // 2
    po.close()


def _remuxed(kind, names):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vorbis_writer as vw
    files = [_load(n) for n in names]
    return vw.chain_files(files) if kind == "chain" else vw.multiplex_files(files)


@pytest.mark.parametrize("kind", ["chain", "multiplex"])
def test_chained_and_multiplexed_logical_streams(golden, kind):
    """f4 (hpp:1433-1484): several logical streams in one physical stream, one after the other or page-interleaved. The
    reference keeps one VorbisStream per live serial; so does the front end here. Every logical stream must come out with
    exactly the descriptors (hence PCM) of the same stream in a file of its own. (The unmodified reference decodes both
    files: 91136 + 15867 frames.)"""
    names = ["stereo44khz", "synth_two_submaps", "mono44khz"]
    po = lib.ParsedOgg(_remuxed(kind, names))
    assert po.n_streams == len(names)
    for i, name in enumerate(names):
        g = golden[name]
        s, b = po.get(i)
        assert s.channels == int(g["channels"]) and s.sample_rate == int(g["sample_rate"])
        assert b.n_packets == len(g["blocksize"])
        pcm, status = ob.synth_batch_raw(s, b, imdct="reference" if ob.reference_lib() is not None else "fast")[:2]
        assert not status.any()
        out = pcm.reshape(int(g["channels"]), -1)
        assert out.shape == g["pcm"].shape and np.abs(out - g["pcm"]).max() <= 1e-5
    po.close()


def test_duplicate_begin_of_stream_is_refused():
    """hpp:1436: a second begin-of-stream page for a serial that is still live."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vorbis_writer as vw
    pages = vw.split_pages(_load("mono44khz"))
    _expect_error(b"".join(pages[:3] + pages[:1] + pages[3:]), "duplicate begin-of-stream")


def _crafted_codebook_file(dim, entries, lookup_type):
    """A ~120-byte Ogg file whose setup header declares one ordered codebook of `entries` entries x `dim` dimensions with a
    VQ lookup and then ends: the size fields are all a decoder gets to see before it would allocate the table."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vorbis_writer as vw
    w = vw.BitWriter()
    w.put(0, 8)                                  # one codebook
    w.put(0x564342, 24); w.put(dim, 16); w.put(entries, 24)
    w.put(1, 1)                                  # ordered
    w.put(0, 5)                                  # codeword length 1 ...
    w.put(entries, vw.ilog(entries))             # ... for every entry
    w.put(lookup_type, 4)
    w.put(0, 32); w.put(0, 32); w.put(0, 4); w.put(0, 1)      # min, delta, value_bits - 1, sequence_p
    ident = b"\x01vorbis" + struct.pack("<IBIiii", 0, 1, 44100, 0, 0, 0) + bytes([8 | (11 << 4), 1])
    comment = b"\x03vorbis" + struct.pack("<I", 0) + struct.pack("<I", 0) + b"\x01"
    return (vw.ogg_page([ident], 7, 0, 0, bos=True) + vw.ogg_page([comment], 7, 1, 0) +
            vw.ogg_page([b"\x05vorbis" + w.bytes()], 7, 2, 0, eos=True))


@pytest.mark.parametrize("dim,entries,lookup_type", [(32768, 1 << 17, 2), (65535, 1 << 20, 1), (65535, (1 << 24) - 1, 2)])
def test_crafted_codebook_sizes_are_refused_not_allocated(dim, entries, lookup_type):
    """Round-1 review: n_entries * dim wrapped in 32 bits (lookup type 2: an empty multiplicand table read out of bounds) or
    asked for 275 GB (lookup type 1: std::bad_alloc across the C boundary). Both must come back as an ordinary stream
    error, from the plain parse entry point and from the parse with raw packets."""
    data = _crafted_codebook_file(dim, entries, lookup_type)
    assert len(data) < 200
    for raw in (False, True):
        with pytest.raises(lib.PovError) as ei:
            lib.ParsedOgg(data, raw_packets=raw)
        assert ei.value.code in (abi.POV_ERR_STREAM, abi.POV_ERR_UNSUPPORTED), ei.value.msg
        assert "codebook" in ei.value.msg, ei.value.msg
