#!/usr/bin/env python3
"""Synthetic Ogg/Vorbis fixtures for what the two bundled files do not cover (SURVEY.md §8c "parity unpinned by data"):
residue type 0 and 1 next to type 2, several submaps, three coupling steps on 5.1, block sizes other than 256/2048,
floor multipliers 1..4, every long/short window transition, sparse / ordered / long-codeword codebooks, lookup type 2 and
sequence_p VQ tables, a codebook with more than 65536 entries (32-bit entry numbers), packets cut short.

    python tests/golden/make_synthetic_golden.py          (build container only: needs oracle/_ref, i.e. /root/reference)

Each scenario is written by tools/vorbis_writer.py (a spec-following bitstream WRITER, no decoder) to
tests/golden/synth_<name>.ogg, decoded by the UNMODIFIED reference (oracle/_ref/ours.bin --debug_out) and stored as
tests/golden/synth_<name>.npz in the layout of make_golden.py. Scenario "shared_submap" is decoded by the hook-patched
libvorbis 1.3.6 instead (oracle/_ref/libvorbis-standalone.bin): with several channels in one type-0/1 residue the
reference advances its partition counter per channel (src/ParseOggVorbis.hpp:755) and is not a usable oracle there.
Deterministic: fixed seeds, so re-running reproduces the committed bytes.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import vorbis_writer as vw  # noqa: E402
from oracle.dumpfile import load_dump  # noqa: E402
from parseoggvorbis_b200 import workloads  # noqa: E402
from tests.golden.make_golden import pack  # noqa: E402

OURS = os.path.join(ROOT, "oracle", "_ref", "ours.bin")
LIBVORBIS = os.path.join(ROOT, "oracle", "_ref", "libvorbis-standalone.bin")


def vq_book(rng, dim, values, skew=0.0, sparse=False, lookup_type=1, sequence_p=False, scale=1.0):
    """A VQ codebook over `values` levels per dimension centred on zero (levels are small integers times `scale`)."""
    n = values ** dim
    if sparse:
        used = rng.random(n) < 0.7
        used[:2] = True
        lens_used = vw.full_tree_lengths(int(used.sum()), rng, skew)
        lengths = [0] * n
        it = iter(lens_used)
        for i in range(n):
            if used[i]:
                lengths[i] = next(it)
    else:
        lengths = vw.full_tree_lengths(n, rng, skew)
    if lookup_type == 1:
        mult = list(range(values))
    else:
        mult = [int(x) for x in rng.integers(0, values, size=n * dim)]
    if sequence_p:          # keep the running sums small
        return vw.Book(dim, lengths, lookup_type, minimum=0.0, delta=scale * 0.25, value_bits=max(1, vw.ilog(values - 1)),
                       sequence_p=True, multiplicands=mult)
    return vw.Book(dim, lengths, lookup_type, minimum=-scale * (values - 1) / 2.0, delta=scale,
                   value_bits=max(1, vw.ilog(values - 1)), sequence_p=False, multiplicands=mult)


def scalar_book(rng, n, skew=0.0):
    return vw.Book(1, vw.full_tree_lengths(n, rng, skew))


def make_floor(rng, books, xs, multiplier, with_subclasses):
    """floor1 over the given X list; the Y books are appended to `books`."""
    R = workloads.FLOOR_RANGE[multiplier]
    tail = list(xs[2:])
    dims = []
    left = len(tail)
    while left:
        d = int(min(left, rng.integers(1, 9)))
        dims.append(d)
        left -= d
    ybook = len(books)
    books.append(scalar_book(rng, R))
    classes = []
    for d in sorted(set(dims)):
        if with_subclasses and d <= 3:
            master = len(books)
            books.append(scalar_book(rng, 1 << d))          # subclass_bits = 1 per dimension
            classes.append((d, vw.FloorClass(d, 1, master, [-1, ybook])))
        else:
            classes.append((d, vw.FloorClass(d, 0, 0, [ybook])))
    class_of_dim = {d: i for i, (d, _) in enumerate(classes)}
    rangebits = vw.ilog(xs[1]) - 1
    return vw.Floor1([class_of_dim[d] for d in dims], [c for _, c in classes], multiplier, rangebits, tail)


def make_residue(rng, books, rtype, begin, end, psize, n_class, vq_ids, cw=2):
    """Residue whose classes use the VQ books `vq_ids` in random passes; class 0 codes nothing."""
    classbook = len(books)
    books.append(vw.Book(cw, vw.full_tree_lengths(n_class ** cw, rng)))
    rows = [[-1] * 8]
    for _ in range(1, n_class):
        row = [-1] * 8
        for p in sorted(rng.choice(8, size=int(rng.integers(1, 4)), replace=False)):
            cand = [b for b in vq_ids if psize % books[b].dim == 0]
            row[int(p)] = int(rng.choice(cand))
        rows.append(row)
    return vw.Residue(rtype, begin, end, psize, classbook, rows)


def choose_packet(rng, s, mode, prev_flag, next_flag, p_unused=0.1, dense=0.5, truncate=None):
    mp = s.mappings[s.modes[mode].mapping]
    n = s.blocksize[s.modes[mode].blockflag]
    draw = rng.random(s.channels) >= p_unused
    if not workloads.closed_under_propagation(draw[None, :], mp.couplings)[0]:
        draw[:] = True
    ys = []
    for c in range(s.channels):
        fl = s.floors[mp.submap_floor[mp.mux[c]]]
        ys.append([int(v) for v in workloads.gen_ys(rng, 1, fl.xs, fl.multiplier)[0]] if draw[c] else None)
    residue = []
    for sm in range(len(mp.submap_floor)):
        chs = [c for c in range(s.channels) if mp.mux[c] == sm]
        r = s.residues[mp.submap_residue[sm]]
        nch, dl = (1, len(chs) * (n // 2)) if r.type == 2 else (len(chs), n // 2)
        _, _, parts = vw.residue_geometry(r, dl)
        cls = np.where(rng.random((nch, parts)) < dense, rng.integers(1, r.n_class, size=(nch, parts)), 0)
        entries = {}
        for pass_ in range(8):
            for p in range(parts):
                for j in range(nch):
                    b = r.books[int(cls[j][p])][pass_]
                    if b >= 0:
                        bk = s.books[b]
                        used = bk.used_entries()
                        entries[(pass_, p, j)] = [used[int(k)] for k in rng.integers(0, len(used), size=r.partition_size // bk.dim)]
        residue.append((cls, entries))
    return vw.PacketChoice(mode, prev_flag, next_flag, ys, residue, truncate)


def block_pattern(rng, P, long_only=False):
    """Block flags with every transition (s->s, s->l, l->s, l->l) present."""
    if long_only:
        return [1] * P
    pat = [0, 0, 1, 1, 0, 1, 0, 0, 0, 1, 1, 1]
    while len(pat) < P:
        pat.append(int(rng.random() < 0.7))
    return pat[:P]


def stream_packets(rng, s, flags, modes_of, **kw):
    out = []
    for i, f in enumerate(flags):
        prev_f = flags[i - 1] if i else 1
        next_f = flags[i + 1] if i + 1 < len(flags) else 1
        out.append(choose_packet(rng, s, int(rng.choice(modes_of[f])), prev_f, next_f, **kw))
    return out


# ---- scenarios ---------------------------------------------------------------------------------------------------
def scen_res0_mono(rng):
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 1, False)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 3, True)
    vq = [len(books) + i for i in range(3)]
    books += [vq_book(rng, 2, 5), vq_book(rng, 4, 3), vq_book(rng, 8, 2, sparse=True)]
    r0 = make_residue(rng, books, 0, 0, 96, 16, 4, vq)
    r1 = make_residue(rng, books, 0, 16, 800, 32, 5, vq)
    s = vw.StreamSetup(1, 16000, (256, 2048), books, [f0, f1], [r0, r1],
                       [vw.Mapping([0], [0], [0]), vw.Mapping([0], [1], [1])], [vw.Mode(0, 0), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 40), {0: [0], 1: [1]}), 17


def scen_two_submaps(rng):
    """Stereo, each channel in its own submap: residue type 1 (ch 0) and type 0 (ch 1), one coupling step, 512/1024."""
    books = []
    f0 = make_floor(rng, books, workloads.scaled_xs(workloads.FIXTURE_XS_SHORT, 128, 256), 2, True)
    f1 = make_floor(rng, books, workloads.scaled_xs(workloads.FIXTURE_XS_LONG, 1024, 512), 4, False)
    vq = [len(books) + i for i in range(4)]
    books += [vq_book(rng, 2, 7, skew=0.8), vq_book(rng, 4, 3), vq_book(rng, 1, 9), vq_book(rng, 2, 4, lookup_type=2)]
    ra = make_residue(rng, books, 1, 0, 400, 16, 6, vq)
    rb = make_residue(rng, books, 0, 8, 300, 8, 3, vq, cw=3)
    maps = [vw.Mapping([0, 1], [0, 0], [0, 1], [(0, 1)]), vw.Mapping([0, 1], [1, 1], [0, 1], [(0, 1)])]
    s = vw.StreamSetup(2, 22050, (512, 1024), books, [f0, f1], [ra, rb], maps, [vw.Mode(0, 0), vw.Mode(1, 1), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 40), {0: [0], 1: [1, 2]}), 5


def scen_surround51(rng):
    """Config 3's shape: 6 channels, long blocks only, LFE alone in submap 1, coupling steps [(0,2),(3,4),(0,1)]."""
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 4, False)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 2, True)
    vq = [len(books) + i for i in range(3)]
    books += [vq_book(rng, 2, 5), vq_book(rng, 4, 3, sparse=True), vq_book(rng, 8, 2)]
    r_main = make_residue(rng, books, 2, 0, 4000, 32, 6, vq)
    r_lfe = make_residue(rng, books, 1, 0, 128, 16, 3, vq)
    mux = [0, 0, 0, 0, 0, 1]
    cp = [(0, 2), (3, 4), (0, 1)]
    maps = [vw.Mapping(mux, [0, 0], [0, 1], cp), vw.Mapping(mux, [1, 1], [0, 1], cp)]
    s = vw.StreamSetup(6, 48000, (256, 2048), books, [f0, f1], [r_main, r_lfe], maps, [vw.Mode(0, 0), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 14, long_only=True), {0: [0], 1: [1]}, p_unused=0.15), 0


def scen_codebooks(rng):
    """Stereo type 2 with awkward codebooks: 24-bit codewords, sparse books, lookup type 2, sequence_p, and an ordered book
    of 2^17 entries (entry numbers need 32 bits in the descriptors)."""
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 4, True)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 2, True)
    vq = [len(books) + i for i in range(5)]
    big = vw.Book(2, [17] * (1 << 17), 1, minimum=-0.5, delta=1.0 / 512.0, value_bits=9, multiplicands=list(range(362)), ordered=True)
    books += [vq_book(rng, 2, 9, skew=0.95), vq_book(rng, 4, 3, sparse=True, skew=0.5), vq_book(rng, 2, 6, lookup_type=2),
              vq_book(rng, 4, 4, sequence_p=True), big]
    r = make_residue(rng, books, 2, 0, 1600, 32, 10, vq)
    maps = [vw.Mapping([0, 0], [0], [0], [(0, 1)]), vw.Mapping([0, 0], [1], [0], [(0, 1)])]
    s = vw.StreamSetup(2, 44100, (256, 2048), books, [f0, f1], [r], maps, [vw.Mode(0, 0), vw.Mode(1, 1)])
    pk = stream_packets(rng, s, block_pattern(rng, 24), {0: [0], 1: [1]})
    # three packets cut short: inside the residue, inside the floor, right after the mode bits (Utils.hpp:389-392 zero fill)
    for i, frac in ((5, 0.6), (9, 0.15), (13, 0.0)):
        full = len(vw.write_audio_packet(s, pk[i]))
        pk[i].truncate_bytes = max(1, int(full * frac))
    return s, pk, 3


def scen_shared_submap(rng):
    """Stereo, both channels in ONE submap with residue type 1 on long blocks and type 0 on short ones."""
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 4, False)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 2, False)
    vq = [len(books) + i for i in range(3)]
    books += [vq_book(rng, 2, 5), vq_book(rng, 4, 3), vq_book(rng, 8, 2)]
    r0 = make_residue(rng, books, 0, 0, 112, 16, 4, vq)
    r1 = make_residue(rng, books, 1, 0, 832, 32, 5, vq)
    maps = [vw.Mapping([0, 0], [0], [0], [(0, 1)]), vw.Mapping([0, 0], [1], [1], [(0, 1)])]
    s = vw.StreamSetup(2, 44100, (256, 2048), books, [f0, f1], [r0, r1], maps, [vw.Mode(0, 0), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 30), {0: [0], 1: [1]}), 0


def scen_residue_edges(rng):
    """Mono: residue `end` beyond the decode length on both block sizes (hpp:697-698 clamps), `begin` beyond it on the short
    one (nothing to read, hpp:704-705), a one-partition residue, multiplier 3 and 1."""
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 3, True)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 1, False)
    vq = [len(books) + i for i in range(2)]
    books += [vq_book(rng, 2, 5), vq_book(rng, 4, 3)]
    r0 = make_residue(rng, books, 1, 200, 5000, 16, 3, vq)       # begin 200 > 128 = n/2: empty
    r1 = make_residue(rng, books, 0, 0, 5000, 1024, 3, vq)       # end 5000 > 1024: one partition of 1024
    s = vw.StreamSetup(1, 8000, (256, 2048), books, [f0, f1], [r0, r1],
                       [vw.Mapping([0], [0], [0]), vw.Mapping([0], [1], [1])], [vw.Mode(0, 0), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 20), {0: [0], 1: [1]}), 40


def scen_bad_vq_book(rng):
    """A residue class points at a scalar (lookup type 0) codebook: the reference fails its decodeVector CHECK
    (hpp:369-370 via :739/:748) on the first packet that uses the class. Both decoders must refuse the file."""
    books = []
    f0 = make_floor(rng, books, workloads.FIXTURE_XS_SHORT, 4, False)
    f1 = make_floor(rng, books, workloads.FIXTURE_XS_LONG, 2, False)
    good = len(books)
    books.append(vq_book(rng, 2, 5))
    bad = len(books)
    books.append(vw.Book(2, vw.full_tree_lengths(25, rng)))      # lookup type 0, same geometry
    classbook = len(books)
    books.append(vw.Book(2, vw.full_tree_lengths(9, rng)))
    r = vw.Residue(1, 0, 512, 32, classbook, [[-1] * 8, [good] + [-1] * 7, [-1, bad] + [-1] * 6])
    s = vw.StreamSetup(1, 44100, (256, 2048), books, [f0, f1], [r],
                       [vw.Mapping([0], [0], [0]), vw.Mapping([0], [1], [0])], [vw.Mode(0, 0), vw.Mode(1, 1)])
    return s, stream_packets(rng, s, block_pattern(rng, 8), {0: [0], 1: [1]}, dense=0.9), 0


SCENARIOS = {
    "res0_mono": (scen_res0_mono, 11, "reference"),
    "two_submaps": (scen_two_submaps, 12, "reference"),
    "surround51": (scen_surround51, 13, "reference"),
    "codebooks": (scen_codebooks, 14, "reference"),
    "shared_submap": (scen_shared_submap, 15, "libvorbis"),
    "residue_edges": (scen_residue_edges, 16, "reference"),
    "bad_vq_book": (scen_bad_vq_book, 17, "reference-fails"),
}


def build(name):
    fn, seed, _ = SCENARIOS[name]
    rng = np.random.default_rng(seed)
    s, packets, trim = fn(rng)
    return s, packets, vw.write_stream(s, packets, serial=0x5000 + seed, packets_per_page=6, trim_last=trim)


def main():
    assert os.path.exists(OURS), "run `make -C oracle ref` first"
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name, (_, _, decoder) in SCENARIOS.items():
            if only and name not in only:
                continue
            s, packets, data = build(name)
            ogg = os.path.join(HERE, "synth_%s.ogg" % name)
            with open(ogg, "wb") as f:
                f.write(data)
            dbg = os.path.join(tmp, name + ".dbg")
            if decoder == "reference-fails":
                r = subprocess.run([OURS, "--in", ogg, "--debug_out", dbg], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                assert r.returncode != 0 and "check failed" in r.stdout, (r.returncode, r.stdout[-300:])
                print("%-14s the reference refuses it: %s" % (name, r.stdout.strip().splitlines()[-1][:120]))
                continue
            if decoder == "reference":
                subprocess.check_call([OURS, "--in", ogg, "--debug_out", dbg], stdout=subprocess.DEVNULL)
            else:
                subprocess.check_call([LIBVORBIS, "--in", ogg, "--debug_out", dbg], stdout=subprocess.DEVNULL)
            dump = load_dump(dbg)
            arrs = pack(dump)
            # what the writer meant is what the decoder read: coded Ys of every used channel
            P = len(packets)
            assert len(dump.packets) == P, (len(dump.packets), P)
            checked = 0
            for pi, pc in enumerate(packets):
                if pc.truncate_bytes is not None:
                    continue
                for c, ys in enumerate(pc.ys):
                    assert arrs["floor_used"][pi, c] == (ys is not None), (name, pi, c)
                    if ys is not None:
                        assert list(arrs["ys"][pi, c, :len(ys)]) == ys, (name, pi, c)
                        checked += 1
            path = os.path.join(HERE, "synth_%s.npz" % name)
            np.savez_compressed(path, **arrs)
            print("%-14s %s: %d packets, %d bytes of Ogg, %d frames, peak %.3f, %d Y lists verified -> %d bytes" % (
                name, decoder, P, len(data), arrs["pcm"].shape[1], float(np.abs(arrs["pcm"]).max()), checked, os.path.getsize(path)))


if __name__ == "__main__":
    main()
