"""GPU tests of the whole-file path: host front-end -> descriptors -> residue kernel -> fused/staged kernels.
Config 1 of BASELINE.json: the bundled fixtures decoded end to end and compared with the reference's dump."""
import ctypes as C
import os

import numpy as np
import pytest

from parseoggvorbis_b200 import abi, lib
from tests import oracle_binding as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = {"stereo44khz": "test.stereo44khz.ogg", "mono44khz": "test.mono44khz.ogg"}
# synthetic streams (tools/vorbis_writer.py) decoded by the unmodified reference / by libvorbis: residue types 0 and 1,
# several submaps, 5.1 with three coupling steps, 512/1024 blocks, awkward codebooks, truncated packets
from tests.conftest import SYNTHETIC  # noqa: E402
FIX.update({n: n + ".ogg" for n in SYNTHETIC})


def _load(name):
    with open(os.path.join(ROOT, "tests", "golden", FIX[name]), "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def ctx():
    c = lib.SynthContext(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(FIX))
def test_decode_fixture_pcm(ctx, golden, name):
    g = golden[name]
    pcm, rate, npk = ctx.decode_ogg(_load(name))
    assert rate == int(g["sample_rate"]) and npk == len(g["blocksize"])
    assert pcm.shape == g["pcm"].shape
    assert np.abs(pcm - g["pcm"]).max() <= 1e-5
    assert ob.snr_db(pcm, g["pcm"]) >= 120.0


@pytest.mark.parametrize("name", list(FIX))
def test_debug_dump_field_by_field(ctx, golden, name, tmp_path):
    """The dump written from the staged device path, parsed back and compared with the reference's dump:
    integer fields and after_residue/after_envelope bit-exact, IMDCT output and PCM <= 1e-5 / >= 120 dB."""
    from oracle.dumpfile import load_dump
    g = golden[name]
    path = str(tmp_path / (name + ".dbg"))
    pcm, _, _ = ctx.decode_ogg(_load(name), debug_out=path)
    keep = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(keep):      # bring it home for the reference's own compare-debug-out.py
        import shutil
        shutil.copyfile(path, os.path.join(keep, name + "_b200.dbg"))
    d = load_dump(path)
    Cn = int(g["channels"])
    assert d.num_channels == Cn and d.sample_rate == int(g["sample_rate"])
    assert d.floor_multipliers == [int(x) for x in g["floor_multipliers"]]
    for i, xs in enumerate(d.floor_xs):
        assert np.array_equal(xs, g["floor_xs"][i, :len(xs)])
    assert len(d.packets) == len(g["blocksize"])
    table = np.load(os.path.join(ROOT, "tests", "golden", "inverse_db_table.npy"))
    oh = of = 0
    worst = 0.0
    for p, pk in enumerate(d.packets):
        n = int(g["blocksize"][p])
        if bool(g.get("has_floor_stages", True)):     # (libvorbis' hooks report positions with another convention)
            assert pk.abs_total_pos == int(g["abs_total_pos"][p])
            assert pk.expected_ending_total_pos == int(g["expected_ending_total_pos"][p])
        for c in range(Cn):
            f = pk.floors[c]
            assert f.floor_number == int(g["floor_number"][p, c])
            if g["floor_used"][p, c]:
                k = int(g["floor_nposts"][f.floor_number])
                assert np.array_equal(f.ys, g["ys"][p, c, :k])
                if bool(g.get("has_floor_stages", True)):      # (a libvorbis-decoded golden holds the coded Ys only)
                    assert np.array_equal(f.final_ys, g["final_ys"][p, c, :k])
                    assert np.array_equal(f.step2_flag, g["step2_flag"][p, c, :k])
                    assert np.array_equal(f.floor, g["floor"][of:of + n])
                    assert np.array_equal(f.floor_outputs.view(np.uint32), table[g["floor"][of:of + n]].view(np.uint32))
            else:
                assert f.ys is None
            assert np.array_equal(pk.after_residue[c], g["after_residue"][oh:oh + n // 2]), (p, c)
            assert np.array_equal(pk.after_envelope[c], g["after_envelope"][oh:oh + n // 2]), (p, c)
            worst = max(worst, float(np.abs(pk.pcm_after_mdct[c] - g["pcm_after_mdct"][of:of + n]).max()))
            oh += n // 2
            of += n
    assert worst <= 1e-5
    out = d.pcm_concat()
    assert np.array_equal(out, pcm)
    assert np.abs(out - g["pcm"]).max() <= 1e-5 and ob.snr_db(out, g["pcm"]) >= 120.0


def test_corpus_decode_matches_single_file(ctx, golden):
    """Config 5 in miniature: replicated fixtures through the threaded front-end, sharded into chunks."""
    files = []
    for i in range(150):
        files.append(_load("stereo44khz" if i % 3 else "mono44khz"))
    frames, total, chk = ctx.decode_corpus(files, host_threads=4)
    exp_frames = [golden["stereo44khz" if i % 3 else "mono44khz"]["pcm"].shape[1] for i in range(150)]
    assert list(frames) == exp_frames
    exp_total = sum(golden["stereo44khz" if i % 3 else "mono44khz"]["pcm"].size for i in range(150))
    assert total == exp_total
    exp_chk = sum(float(golden["stereo44khz" if i % 3 else "mono44khz"]["pcm"].astype(np.float64).sum()) for i in range(150))
    assert abs(chk - exp_chk) <= 1e-3 * max(1.0, abs(exp_chk)) + 0.05


def test_corpus_decode_reuses_its_state_and_survives_a_bad_file(ctx, golden):
    """pov_decode_corpus keeps its sibling context, slots and pinned staging pool on the context: a second call with
    another thread count gives the same answer, a corrupt file is reported with the reference's message shape, and the
    context keeps working afterwards."""
    good = [_load("stereo44khz")] * 200
    f1, t1, c1 = ctx.decode_corpus(good, host_threads=3)
    f2, t2, c2 = ctx.decode_corpus(good, host_threads=16)       # more threads than chunks
    assert list(f1) == list(f2) and t1 == t2 and abs(c1 - c2) <= 1e-6 * max(1.0, abs(c1))
    assert t1 == 200 * golden["stereo44khz"]["pcm"].size
    bad = bytearray(good[0]); bad[9000] ^= 0xFF                   # inside an audio page: the page CRC no longer matches
    files = list(good[:100]) + [bytes(bad)] + list(good[:100])
    with pytest.raises(RuntimeError) as ei:
        ctx.decode_corpus(files, host_threads=8)
    assert "file 100" in str(ei.value) and "check failed" in str(ei.value)
    f3, t3, c3 = ctx.decode_corpus(good, host_threads=5)
    assert list(f3) == list(f1) and t3 == t1 and abs(c3 - c1) <= 1e-6 * max(1.0, abs(c1))


def test_reference_shaped_entry_point():
    L = lib.load()
    data = _load("mono44khz")
    err = C.c_char_p(None)
    assert L.pov_ogg_vorbis_full_read_from_memory(data, len(data), C.byref(err)) == 0
    bad = bytearray(data); bad[3000] ^= 0xFF
    assert L.pov_ogg_vorbis_full_read_from_memory(bytes(bad), len(bad), C.byref(err)) == 1
    assert b"check failed" in err.value


def test_entries_batch_staged_equals_fused(ctx):
    """POV_INPUT_ENTRIES batches straight from the host parser: residue kernel + fused == residue kernel + staged."""
    po = lib.ParsedOgg(_load("stereo44khz"))
    s, b = po.get(0)
    sid = C.c_uint32(0)
    ctx._check(ctx.L.pov_setup_register(ctx.ctx, C.byref(s), C.byref(sid)))
    st = abi.pov_stream.from_address(C.addressof(b.streams.contents))
    st.setup_id = sid.value
    h = C.c_void_p(None)
    ctx._check(ctx.L.pov_batch_upload(ctx.ctx, C.byref(b), C.byref(h)))
    out = []
    for run in (ctx.L.pov_batch_run, ctx.L.pov_batch_run_staged):
        ctx._check(run(ctx.ctx, h))
        pcm = np.empty(int(b.pcm_floats), np.float32)
        ctx._check(ctx.L.pov_batch_fetch_pcm(ctx.ctx, h, pcm.ctypes.data_as(C.POINTER(C.c_float)), pcm.size, 1))
        out.append(pcm)
    assert np.abs(out[0] - out[1]).max() <= 1e-5 and ob.snr_db(out[0], out[1]) >= 120.0
    st.setup_id = 0
    ctx.L.pov_batch_free(ctx.ctx, h)
    po.close()


def test_corpus_output_edge_delivers_every_sample_in_file_order(ctx, golden):
    """pov_decode_corpus_pcm: the sink sees the PCM of 150 mixed files (bundled + synthetic streams, several chunks), in
    file order, sample for sample what the reference decodes (hpp:966-973, 1047-1053)."""
    names = ["stereo44khz", "mono44khz", "synth_res0_mono", "synth_two_submaps", "synth_surround51"]
    files = [_load(names[i % len(names)]) for i in range(150)]
    seen = []

    bad = []

    def sink(file_index, pcm):          # (an exception inside a ctypes callback would be swallowed: record, assert afterwards)
        ref = golden[names[file_index % len(names)]]["pcm"]
        if pcm.shape != ref.shape or float(np.abs(pcm - ref).max()) > 1e-5:
            bad.append(file_index)
        seen.append(file_index)
        return False
    frames, total, chk = ctx.decode_corpus_pcm(files, sink, host_threads=4)
    assert seen == list(range(150)) and not bad, bad
    assert list(frames) == [golden[names[i % len(names)]]["pcm"].shape[1] for i in range(150)]
    assert total == sum(golden[names[i % len(names)]]["pcm"].size for i in range(150))


def test_corpus_output_edge_stops_when_the_sink_says_so(ctx):
    files = [_load("mono44khz")] * 200
    calls = []

    def sink(file_index, pcm):
        calls.append(file_index)
        return file_index == 70
    with pytest.raises(RuntimeError) as ei:
        ctx.decode_corpus_pcm(files, sink, host_threads=3)
    assert "gotPcmData" in str(ei.value) and calls[-1] == 70 and calls == list(range(71))
    f, t, _ = ctx.decode_corpus(files, host_threads=3)           # the context keeps working
    assert len(f) == 200 and t > 0


@pytest.mark.parametrize("kind", ["chain", "multiplex"])
def test_chained_and_multiplexed_files_through_the_corpus_decode(ctx, golden, kind):
    """f4 (hpp:1433-1484): files holding several logical streams (chained / page-interleaved, different channel counts and
    rates). The sink is called once per logical stream, in begin-of-stream order, with the PCM the same stream gives in a
    file of its own; frames_out holds the file's total. The unmodified reference decodes the same files."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vorbis_writer as vw
    names = ["stereo44khz", "synth_two_submaps", "mono44khz"]
    mux = vw.chain_files if kind == "chain" else vw.multiplex_files
    files = [mux([_load(n) for n in names], first_serial=0x100 * (i + 1)) if i % 2 == 0 else _load("mono44khz") for i in range(40)]
    calls, bad = [], []

    def sink(file_index, pcm):
        k = sum(1 for f in calls if f == file_index)
        ref = golden[names[k] if file_index % 2 == 0 else "mono44khz"]["pcm"]
        if pcm.shape != ref.shape or float(np.abs(pcm - ref).max()) > 1e-5:
            bad.append((file_index, k))
        calls.append(file_index)
        return False
    frames, total, _ = ctx.decode_corpus_pcm(files, sink, host_threads=3)
    assert not bad, bad
    assert calls == [i for i in range(40) for _ in range(3 if i % 2 == 0 else 1)]
    per_file = sum(golden[n]["pcm"].shape[1] for n in names)
    assert list(frames) == [per_file if i % 2 == 0 else golden["mono44khz"]["pcm"].shape[1] for i in range(40)]


def test_chained_streams_of_one_layout_decode_to_one_pcm_array(ctx, golden):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import vorbis_writer as vw
    a = _load("stereo44khz")
    pcm, rate, npk = ctx.decode_ogg(vw.chain_files([a, a]))
    ref = np.concatenate([golden["stereo44khz"]["pcm"]] * 2, axis=1)
    assert rate == 44100 and npk == 2 * len(golden["stereo44khz"]["blocksize"])
    assert pcm.shape == ref.shape and np.abs(pcm - ref).max() <= 1e-5


def _stage_all(ctx, h, b, s, npk_blocksizes, stage, count_of):
    out = []
    for p, n in enumerate(npk_blocksizes):
        for c in range(int(s.channels)):
            k = count_of(int(n))
            buf = np.empty(k, np.float32)
            ctx._check(ctx.L.pov_batch_fetch_stage(ctx.ctx, h, p, c, stage, buf.ctypes.data_as(C.c_void_p), buf.nbytes))
            out.append(buf)
    return out


@pytest.mark.parametrize("name", list(FIX))
def test_device_entropy_decode_is_bit_exact(ctx, golden, name):
    """f2: the audio packets go to the device as raw bytes (POV_INPUT_PACKETS) and k_packet_decode walks them: floor flags,
    coded Ys, classifications and VQ entry numbers. `after_residue` must equal the reference dump bit for bit (hpp:696-760,
    347-374), the status words must be clean, and the PCM must equal what the host-decoded batch gives, bit for bit."""
    g = golden[name]
    data = _load(name)
    po = lib.ParsedOgg(data, raw_packets=True)
    s, b = po.get(0)
    assert b.input_kind == abi.POV_INPUT_PACKETS
    sid = C.c_uint32(0)
    ctx._check(ctx.L.pov_setup_register(ctx.ctx, C.byref(s), C.byref(sid)))
    st = abi.pov_stream.from_address(C.addressof(b.streams.contents))
    st.setup_id = sid.value
    h = C.c_void_p(None)
    ctx._check(ctx.L.pov_batch_upload(ctx.ctx, C.byref(b), C.byref(h)))
    ctx._check(ctx.L.pov_batch_run_staged(ctx.ctx, h))
    status = np.zeros(int(b.n_packets), np.uint32)
    ctx._check(ctx.L.pov_batch_status(ctx.ctx, h, status.ctypes.data_as(C.POINTER(C.c_uint32)), status.size))
    assert not status.any()
    res = _stage_all(ctx, h, b, s, g["blocksize"], abi.POV_STAGE_AFTER_RESIDUE, lambda n: n // 2)
    got = np.concatenate(res)
    assert np.array_equal(got, g["after_residue"])            # (value equality: libvorbis' golden carries a few -0.0)
    env = np.concatenate(_stage_all(ctx, h, b, s, g["blocksize"], abi.POV_STAGE_AFTER_ENVELOPE, lambda n: n // 2))
    assert np.array_equal(env, g["after_envelope"])
    ctx._check(ctx.L.pov_batch_run(ctx.ctx, h))
    pcm_dev = np.empty(int(b.pcm_floats), np.float32)
    ctx._check(ctx.L.pov_batch_fetch_pcm(ctx.ctx, h, pcm_dev.ctypes.data_as(C.POINTER(C.c_float)), pcm_dev.size, 1))
    st.setup_id = 0
    ctx.L.pov_batch_free(ctx.ctx, h)
    po.close()
    # the same file with the host walk
    ctx.set_device_entropy(False)
    try:
        pcm_host, _, _ = ctx.decode_ogg(data)
    finally:
        ctx.set_device_entropy(True)
    pcm_auto, _, _ = ctx.decode_ogg(data)
    assert np.array_equal(pcm_host, pcm_auto)
    assert np.array_equal(pcm_dev.reshape(pcm_host.shape), pcm_host)
    assert np.abs(pcm_auto - g["pcm"]).max() <= 1e-5


@pytest.mark.parametrize("name", ["stereo44khz", "mono44khz", "synth_two_submaps", "synth_surround51"])
@pytest.mark.parametrize("raw_packets", [True, False])
def test_feature_matrices_match_the_reference_readers(ctx, name, raw_packets):
    """f3: the (frames, dim) matrices the reference builds in Python from a debug dump (returnn_import.py:74-115,
    demo_live_extract.py:262-505), produced on the device. Goldens: the reference's own readers on the reference
    decoder's dump (tests/golden/make_feature_golden.py). Kinds without exp() are bit exact."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "features.npz"))
    data = _load(name)
    checked = 0
    for key in gold.files:
        fx, kind, dim = key.split("/")
        if fx != name:
            continue
        ref = gold[key]
        got = ctx.features_from_raw_bytes(data, int(dim), kind, raw_packets=raw_packets)
        assert got.shape == ref.shape, (key, got.shape, ref.shape)
        if kind == "residue_ys_with_floor":
            assert np.allclose(got, ref, rtol=2e-6, atol=1e-9), (key, float(np.abs(got - ref).max()))
        else:
            assert np.array_equal(got, ref), (key, float(np.abs(got - ref).max()))
        checked += 1
    assert checked >= 10


@pytest.mark.parametrize("device_entropy", [True, False])
def test_scalar_book_used_as_vq_book_fails_on_the_device_too(ctx, device_entropy):
    """The same file through the whole-file decode: with the entropy decode on the device the packet is flagged
    POV_PKT_VQ_ENTRY by the kernels and the decode fails with the reference's check (hpp:739,748)."""
    with open(os.path.join(ROOT, "tests", "golden", "synth_bad_vq_book.ogg"), "rb") as f:
        data = f.read()
    ctx.set_device_entropy(device_entropy)
    try:
        with pytest.raises(lib.PovError) as ei:
            ctx.decode_ogg(data)
    finally:
        ctx.set_device_entropy(True)
    assert ei.value.code == abi.POV_ERR_STREAM and "check failed" in ei.value.msg


def test_packets_spanning_pages_on_the_device(ctx, golden):
    """f4 end to end: a stream whose packets continue across pages, decoded with both entropy-decode modes."""
    from tests.test_front_end_cpu import _spanning_twin
    data = _spanning_twin("two_submaps")
    g = golden["synth_two_submaps"]
    with pytest.raises(lib.PovError):
        ctx.decode_ogg(data)                       # default: refused like the reference (hpp:89)
    ctx.set_page_spanning(True)
    try:
        for dev in (True, False):
            ctx.set_device_entropy(dev)
            pcm, _, npk = ctx.decode_ogg(data)
            assert npk == len(g["blocksize"]) and pcm.shape == g["pcm"].shape
            assert np.abs(pcm - g["pcm"]).max() <= 1e-5
    finally:
        ctx.set_page_spanning(False)
        ctx.set_device_entropy(True)


def _flip_bits_in_audio_packets(data: bytes, rng, flips: int) -> bytes:
    """Corrupt `flips` random bits inside the bodies of the audio pages (page 4 onwards) and repair the page CRCs, so that
    the damage reaches the packet decoders instead of being caught by the container check (hpp:92-98)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from vorbis_writer import ogg_crc
    buf = bytearray(data)
    pages, pos = [], 0
    while pos + 27 <= len(buf):
        nseg = buf[pos + 26]
        body = sum(buf[pos + 27:pos + 27 + nseg])
        pages.append((pos, 27 + nseg, body))
        pos += 27 + nseg + body
    audio = pages[3:]
    for _ in range(flips):
        p, hdr, body = audio[int(rng.integers(0, len(audio)))]
        if body == 0:
            continue
        at = p + hdr + int(rng.integers(0, body))
        buf[at] ^= 1 << int(rng.integers(0, 8))
    for p, hdr, body in audio:
        buf[p + 22:p + 26] = b"\0\0\0\0"
        buf[p + 22:p + 26] = ogg_crc(bytes(buf[p:p + hdr + body])).to_bytes(4, "little")
    return bytes(buf)


@pytest.mark.parametrize("name", ["stereo44khz", "synth_two_submaps", "synth_codebooks"])
def test_corrupted_packets_decode_identically_on_host_and_device(ctx, name):
    """Random bit flips inside audio packets: whatever the bits now say, the device walk (k_packet_decode) and the host walk
    must read the same thing — same error or the same PCM, bit for bit. (Zero fill past the packet end, long codewords,
    classification words and cascade order are all exercised by garbage far better than by valid streams.)"""
    rng = np.random.default_rng(99)
    base = _load(name)
    agree_ok = agree_err = 0
    for trial in range(24):
        data = _flip_bits_in_audio_packets(base, rng, flips=int(rng.integers(1, 12)))
        out = []
        for dev in (True, False):
            ctx.set_device_entropy(dev)
            try:
                pcm, _, _ = ctx.decode_ogg(data)
                out.append(pcm)
            except lib.PovError as e:
                out.append(e.code)
            finally:
                ctx.set_device_entropy(True)
        if isinstance(out[0], int) or isinstance(out[1], int):
            assert isinstance(out[0], int) and isinstance(out[1], int), (trial, out[0] if isinstance(out[0], int) else "ok", out[1] if isinstance(out[1], int) else "ok")
            agree_err += 1
        else:
            assert out[0].shape == out[1].shape and np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32)), trial
            agree_ok += 1
    assert agree_ok + agree_err == 24 and agree_ok >= 3


@pytest.mark.parametrize("name", ["stereo44khz", "mono44khz", "synth_res0_mono", "synth_two_submaps", "synth_surround51",
                                  "synth_codebooks", "synth_residue_edges"])
def test_production_kernel_floor_stage_is_bit_exact(ctx, golden, name):
    """The integer floor1 stage of k_warp_synth ITSELF (not of the staged kernels): final Y values and step-2 flags of every
    decoded curve, copied out of the kernel's shared memory by its parity hook, against the reference dump's
    "floor1 final_ys" / "floor1 step2_flag" (hpp:521-559). The dump lists posts in bitstream order, the kernel keeps them in
    ascending-x order (hpp:458-469)."""
    g = golden[name]
    po = lib.ParsedOgg(_load(name), raw_packets=True)
    s, b = po.get(0)
    sid = C.c_uint32(0)
    ctx._check(ctx.L.pov_setup_register(ctx.ctx, C.byref(s), C.byref(sid)))
    st = abi.pov_stream.from_address(C.addressof(b.streams.contents))
    st.setup_id = sid.value
    h = C.c_void_p(None)
    ctx._check(ctx.L.pov_batch_upload(ctx.ctx, C.byref(b), C.byref(h)))
    try:
        assert ctx.L.pov_batch_kernel_name(ctx.ctx, h) == b"k_warp_synth"
        Cn, P = int(g["channels"]), int(b.n_packets)
        rec = np.zeros((P, Cn, 72), np.uint8)
        ctx._check(ctx.L.pov_batch_fetch_fast_floor(ctx.ctx, h, rec.ctypes.data_as(C.POINTER(C.c_uint8)), rec.nbytes))
        checked = 0
        for p in range(P):
            for c in range(Cn):
                if not g["floor_used"][p, c]:
                    continue
                fl = int(g["floor_number"][p, c])
                k = int(g["floor_nposts"][fl])
                order = np.argsort(g["floor_xs"][fl, :k], kind="stable")
                want_y = np.minimum(g["final_ys"][p, c, :k][order], 255).astype(np.uint8)
                want_f = g["step2_flag"][p, c, :k][order]
                mask = int.from_bytes(rec[p, c, 64:72].tobytes(), "little")
                got_f = np.array([(mask >> i) & 1 for i in range(k)], bool)
                assert np.array_equal(rec[p, c, :k], want_y), (p, c)
                assert np.array_equal(got_f, want_f), (p, c)
                checked += 1
        assert checked > 10
    finally:
        st.setup_id = 0
        ctx.L.pov_batch_free(ctx.ctx, h)
        po.close()
