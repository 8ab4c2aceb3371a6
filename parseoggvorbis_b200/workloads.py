"""Synthetic descriptor batches for the benchmark configs of SURVEY.md §8(d) and batches rebuilt from the
reference's debug dumps (golden fixtures). Pure numpy input generation — no decode arithmetic lives here.

Config 2: 44.1 kHz stereo, mixed 256/2048, one coupling step, P packets per stream.
Config 3: 48 kHz 5.1, long blocks only, 2 submaps, 3 coupling steps [(0,2),(3,4),(0,1)].
Config 4: 16 kHz mono speech clips, 100 packets per clip.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi

# floor1 X lists of the two bundled fixtures (dump entries "floor1_unpack xs"; SURVEY.md §8 header):
FIXTURE_XS_SHORT = [0, 128, 14, 4, 58, 2, 8, 28, 90]                       # multiplier 4, blocksize 256
FIXTURE_XS_LONG = [0, 1024, 93, 23, 372, 6, 46, 186, 750, 14, 33, 65, 130, 260, 556, 3, 10, 18, 28, 39, 55,
                   79, 111, 158, 220, 312, 464, 650, 850]                   # multiplier 2, blocksize 2048
FLOOR_RANGE = {1: 256, 2: 128, 3: 86, 4: 64}


def scaled_xs(xs: Sequence[int], half_from: int, half_to: int) -> List[int]:
    """Rescale an X list made for n/2 = half_from to n/2 = half_to, keeping the values distinct."""
    out = [0, half_to]
    seen = {0, half_to}
    for x in xs[2:]:
        v = max(1, min(half_to - 1, (x * half_to) // half_from))
        while v in seen:
            v += 1
        assert v < half_to
        seen.add(v)
        out.append(v)
    return out


def make_setup(channels: int, blocksizes=(256, 2048), sample_rate=44100,
               couplings: Sequence[Tuple[int, int]] = (), mux: Optional[Sequence[int]] = None,
               n_submaps: int = 1, multipliers=(4, 2)) -> abi.Setup:
    """A setup with one floor per blocksize class (the fixtures' post lists, rescaled if needed), one mapping
    per class (same channel routing) and two modes: mode 0 = short, mode 1 = long."""
    bs0, bs1 = blocksizes
    xs0 = FIXTURE_XS_SHORT if bs0 == 256 else scaled_xs(FIXTURE_XS_SHORT, 128, bs0 // 2)
    xs1 = FIXTURE_XS_LONG if bs1 == 2048 else scaled_xs(FIXTURE_XS_LONG, 1024, bs1 // 2)
    floors = [abi.Floor1(xs0, multipliers[0]), abi.Floor1(xs1, multipliers[1])]
    mux = list(mux) if mux is not None else [0] * channels
    mappings = [
        abi.Mapping(mux=mux, submap_floor=[0] * n_submaps, submap_residue=[0] * n_submaps, couplings=list(couplings)),
        abi.Mapping(mux=mux, submap_floor=[1] * n_submaps, submap_residue=[0] * n_submaps, couplings=list(couplings)),
    ]
    modes = [abi.Mode(0, 0), abi.Mode(1, 1)]
    return abi.Setup(channels=channels, sample_rate=sample_rate, blocksize=(bs0, bs1), floors=floors,
                     mappings=mappings, modes=modes)


def block_sequence(P: int, rng: np.random.Generator, p_short: float = 0.05, long_only: bool = False) -> np.ndarray:
    """Block flags (1 = long) of P packets: Markov chain, P(long->short) = p_short, short runs of 2..8."""
    if long_only:
        return np.ones(P, np.uint8)
    flags = np.ones(P, np.uint8)
    i = 0
    while i < P:
        if rng.random() < p_short:
            run = int(rng.integers(2, 9))
            flags[i:i + run] = 0
            i += run
        else:
            i += 1
    return flags


@dataclass
class StreamPlan:
    """Per-packet geometry of one stream (all arrays length P)."""
    blockflag: np.ndarray
    n: np.ndarray            # blocksize per packet
    window_flags: np.ndarray
    emit: np.ndarray
    pcm_off: np.ndarray
    frames: int


def plan_stream(blockflag: np.ndarray, blocksizes, trim_last: int = 0) -> StreamPlan:
    bs = np.asarray(blocksizes, np.int64)
    n = bs[blockflag.astype(np.int64)]
    P = len(n)
    prev_long = np.concatenate([[1], blockflag[:-1]]).astype(np.uint8)
    next_long = np.concatenate([blockflag[1:], [1]]).astype(np.uint8)
    wf = np.where(blockflag == 1, prev_long | (next_long << 1), 0).astype(np.uint8)
    emit = np.zeros(P, np.int64)
    emit[1:] = n[:-1] // 4 + n[1:] // 4
    if trim_last and P > 1:
        emit[-1] = max(0, emit[-1] - trim_last)
    off = np.concatenate([[0], np.cumsum(emit)[:-1]])
    return StreamPlan(blockflag.astype(np.uint8), n, wf, emit, off, int(emit.sum()))


def floor_neighbours(xs: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
    """low / high neighbour of every post among the posts before it (Vorbis I 9.2.4/9.2.5, src/Utils.hpp:60-118)."""
    n = len(xs)
    lo = np.zeros(n, np.int64)
    hi = np.ones(n, np.int64)
    for i in range(2, n):
        below = [j for j in range(i) if xs[j] < xs[i]]
        above = [j for j in range(i) if xs[j] > xs[i]]
        lo[i] = max(below, key=lambda j: xs[j])
        hi[i] = min(above, key=lambda j: xs[j])
    return lo, hi


def gen_ys(rng: np.random.Generator, count: int, xs: Sequence[int], multiplier: int) -> np.ndarray:
    """Coded floor1 Y lists that unwrap to audio-like curves. A target curve is drawn first — the statistics of the two
    bundled fixtures (tests/golden/*.npz): about 0.49 of the range at x = 0 falling to 0.25 at the last bin with a spread
    of 0.06, 65 % of the inner posts left on their prediction — and then ENCODED post by post, i.e. the inverse of the
    amplitude unwrap of hpp:536-557 (what an encoder does). The floor therefore lands where real audio has it
    (inverse-dB values of 1e-6 .. 1e-2) and the PCM stays inside [-1, 1], which lets the tests assert north_star's
    literal 1e-5 bound. Input generation only: nothing here is used to check a result."""
    R = FLOOR_RANGE[multiplier]
    x = np.asarray(xs, np.int64)
    n = len(x)
    lo, hi = floor_neighbours(xs)
    base = (0.49 - 0.24 * np.sqrt(x / float(x[1]))) * R                       # falls quickly, like a spectrum envelope
    target = np.clip(np.rint(base[None, :] + rng.normal(0.0, 0.06 * R, size=(count, n))), 1, R - 1).astype(np.int64)
    keep = rng.random((count, n)) < 0.65                                        # inner posts coded as "on the prediction"
    final = np.zeros((count, n), np.int64)
    coded = np.zeros((count, n), np.int64)
    final[:, :2] = target[:, :2]
    coded[:, :2] = target[:, :2]
    for i in range(2, n):
        y0, y1 = final[:, lo[i]], final[:, hi[i]]
        ady, adx = np.abs(y1 - y0), x[hi[i]] - x[lo[i]]
        off = ady * (x[i] - x[lo[i]]) // adx
        pred = np.where(y1 < y0, y0 - off, y0 + off)
        f = np.where(keep[:, i], pred, target[:, i])
        d = f - pred
        high_room, low_room = R - pred, pred
        room = 2 * np.minimum(high_room, low_room)
        inside = np.where(d > 0, 2 * d, -2 * d - 1)                            # hpp:553-554 inverted
        beyond = np.where(high_room > low_room, d + low_room, -d + high_room - 1)   # hpp:546-551 inverted
        val = np.where(d == 0, 0, np.where(inside < room, inside, beyond))
        coded[:, i] = val
        final[:, i] = f
    assert coded.min() >= 0 and coded.max() < max(R, 2 * R)
    return coded.astype(np.uint16)


def gen_spectra(rng: np.random.Generator, rows: int, half: int) -> np.ndarray:
    """after_residue-like vectors: round(Laplace(0,1.5)) below 0.78*half, zero above."""
    cut = int(0.78 * half)
    out = np.zeros((rows, half), np.float32)
    out[:, :cut] = np.round(rng.laplace(0.0, 1.5, size=(rows, cut))).astype(np.float32)
    return out


def propagated(used: np.ndarray, couplings: Sequence[Tuple[int, int]]) -> np.ndarray:
    """hpp:1174-1180 on a (packets, channels) bool array: a coupling step with one decoded channel marks both."""
    prop = used.copy()
    for m, a in couplings:
        either = prop[:, m] | prop[:, a]
        prop[:, m] = either
        prop[:, a] = either
    return prop


def closed_under_propagation(used: np.ndarray, couplings: Sequence[Tuple[int, int]]) -> np.ndarray:
    """Rows of `used` for which the reference's single pass over the coupling steps already is the fixed point. Where it is
    not (a later step marks a channel that an earlier step would then have spread further), the reference un-couples a
    decoded channel into one that keeps its raw residue — integers of magnitude 10 straight into the PCM. Encoders order
    their steps so that this cannot happen; the audio-like generators avoid such packets."""
    once = propagated(used, couplings)
    return (propagated(once, couplings) == once).all(1)


# A channel whose floor is unused and that no coupling step pulls in keeps its residue vector as it is (the reference skips
# the floor multiplication, hpp:1247). A decoder gets zeros there (the residue of such a channel is not coded); the
# synthetic batches keep the case alive with small non-zero values so that the path is exercised at audio-like amplitude.
QUIET = np.float32(2.0 ** -10)


def build_dense_batch(setup: abi.Setup, plans: Sequence[StreamPlan], rng: np.random.Generator,
                      p_unused: float = 0.01, setup_id: int = 0,
                      pcm_layout: int = abi.POV_PCM_PLANAR) -> abi.Batch:
    """Random dense-spectra batch for the given stream plans (all streams use `setup`)."""
    C = setup.channels
    S = len(plans)
    P = sum(len(p.n) for p in plans)
    packets = np.zeros(P, abi.PACKET_DTYPE)
    streams = np.zeros(S, abi.STREAM_DTYPE)
    posts = [len(setup.floors[setup.mappings[m.mapping].submap_floor[0]].xs) for m in setup.modes]
    ranges = [FLOOR_RANGE[setup.floors[setup.mappings[m.mapping].submap_floor[0]].multiplier] for m in setup.modes]
    first = 0
    pcm_base = 0
    bf_all = np.concatenate([p.blockflag for p in plans])
    n_all = np.concatenate([p.n for p in plans])
    for s, p in enumerate(plans):
        k = len(p.n)
        sl = slice(first, first + k)
        packets["stream"][sl] = s
        packets["mode"][sl] = p.blockflag       # mode 0 = short, mode 1 = long
        packets["window_flags"][sl] = p.window_flags
        packets["emit_frames"][sl] = p.emit
        packets["pcm_off"][sl] = p.pcm_off
        streams[s] = (setup_id, first, k, 0, p.frames, pcm_base)
        pcm_base += p.frames * C
        first += k
    used = rng.random((P, C)) >= p_unused
    for cls in (0, 1):                   # packets whose propagation is not closed get every floor (see closed_under_propagation)
        rows = np.nonzero(bf_all == cls)[0]
        if len(rows):
            bad = ~closed_under_propagation(used[rows], setup.mappings[setup.modes[cls].mapping].couplings)
            used[rows[bad]] = True
    packets["floor_used"] = (used.astype(np.uint16) << np.arange(C, dtype=np.uint16)).sum(1).astype(np.uint16)
    # Y arena: only used channels carry a list
    per_packet_posts = np.where(bf_all == 1, posts[1], posts[0]).astype(np.int64)
    ys_count = used.sum(1) * per_packet_posts
    packets["ys_off"] = np.concatenate([[0], np.cumsum(ys_count)[:-1]])
    ys = np.zeros(int(ys_count.sum()), np.uint16)
    for cls in (0, 1):
        rows = np.nonzero(bf_all == cls)[0]
        if len(rows) == 0:
            continue
        nrows = int(used[rows].sum())
        fl = setup.floors[setup.mappings[setup.modes[cls].mapping].submap_floor[0]]
        vals = gen_ys(rng, nrows, fl.xs, fl.multiplier)
        # destination index of every (packet, used channel) list
        starts = (packets["ys_off"][rows].astype(np.int64)[:, None]
                  + (np.cumsum(used[rows], 1) - used[rows]).astype(np.int64) * posts[cls])[used[rows]]
        idx = (starts[:, None] + np.arange(posts[cls], dtype=np.int64)[None, :]).ravel()
        ys[idx] = vals.ravel()
    # dense spectra arena
    half_all = n_all // 2
    spec_count = half_all * C
    packets["spec_off"] = np.concatenate([[0], np.cumsum(spec_count)[:-1]])
    spec = np.zeros(int(spec_count.sum()), np.float32)
    for cls in (0, 1):
        rows = np.nonzero(bf_all == cls)[0]
        if len(rows) == 0:
            continue
        half = int(setup.blocksize[cls] // 2)
        data = gen_spectra(rng, len(rows) * C, half)
        prop = propagated(used[rows], setup.mappings[setup.modes[cls].mapping].couplings)
        data = (data.reshape(len(rows), C, half) * np.where(prop, np.float32(1), QUIET)[:, :, None]).reshape(len(rows) * C, half)
        idx = (packets["spec_off"][rows].astype(np.int64)[:, None] + np.arange(C * half, dtype=np.int64)[None, :]).ravel()
        spec[idx] = data.ravel()
    return abi.Batch(streams=streams, packets=packets, ys=ys, payload=spec, pcm_floats=pcm_base,
                     input_kind=abi.POV_INPUT_DENSE, pcm_layout=pcm_layout)


def replicate_batch(b: abi.Batch, times: int) -> abi.Batch:
    """`times` copies of a batch as independent streams with their own arena regions (distinct HBM addresses)."""
    S, P = len(b.streams), len(b.packets)
    streams = np.tile(b.streams, times)
    packets = np.tile(b.packets, times)
    rep_s = np.repeat(np.arange(times, dtype=np.uint64), S)
    rep_p = np.repeat(np.arange(times, dtype=np.uint64), P)
    streams["first_packet"] += (rep_s * P).astype(np.uint32)
    streams["pcm_base"] += rep_s * np.uint64(b.pcm_floats)
    packets["stream"] += (rep_p * S).astype(np.uint32)
    packets["ys_off"] += rep_p * np.uint64(len(b.ys))
    packets["spec_off"] += rep_p * np.uint64(b.payload.size if b.input_kind == abi.POV_INPUT_DENSE else b.payload.nbytes)
    return abi.Batch(streams=streams, packets=packets, ys=np.tile(b.ys, times), payload=np.tile(b.payload, times),
                     pcm_floats=b.pcm_floats * times, input_kind=b.input_kind, pcm_layout=b.pcm_layout)


# ---- the named configs ---------------------------------------------------------------------------------------
def config2(P: int = 4096, streams: int = 1, seed: int = 0, distinct: int = 1):
    """Synthetic 44.1 kHz stereo, mixed 256/2048 (BASELINE.json configs[1]). `distinct` independently generated
    streams are replicated up to `streams` streams."""
    rng = np.random.default_rng(seed)
    setup = make_setup(2, (256, 2048), 44100, couplings=[(0, 1)])
    plans = [plan_stream(block_sequence(P, rng), setup.blocksize) for _ in range(distinct)]
    b = build_dense_batch(setup, plans, rng)
    if streams > distinct:
        b = replicate_batch(b, streams // distinct)
    return setup, b


def config3(P: int = 4096, streams: int = 1, seed: int = 1, distinct: int = 1):
    """Synthetic 48 kHz 5.1, long blocks only, 2 submaps (LFE alone), 3 coupling steps; channel 0 is in two steps
    to pin the reverse-order rule (hpp:1214)."""
    rng = np.random.default_rng(seed)
    setup = make_setup(6, (256, 2048), 48000, couplings=[(0, 2), (3, 4), (0, 1)], mux=[0, 0, 0, 0, 0, 1],
                       n_submaps=2)
    plans = [plan_stream(block_sequence(P, rng, long_only=True), setup.blocksize) for _ in range(distinct)]
    b = build_dense_batch(setup, plans, rng)
    if streams > distinct:
        b = replicate_batch(b, streams // distinct)
    return setup, b


def config4(clips: int = 100, packets_per_clip: int = 100, seed: int = 2, blocksizes=(256, 2048), replicate: int = 1):
    """16 kHz mono speech clips (returnn_import use case): `clips` independent streams of 100 packets."""
    rng = np.random.default_rng(seed)
    setup = make_setup(1, blocksizes, 16000)
    plans = [plan_stream(block_sequence(packets_per_clip, rng), setup.blocksize) for _ in range(clips)]
    b = build_dense_batch(setup, plans, rng)
    if replicate > 1:
        b = replicate_batch(b, replicate)
    return setup, b


# ---- batches rebuilt from the reference's own dump (golden fixtures) -------------------------------------------
def golden_setup_and_batch(g: Dict[str, np.ndarray]) -> Tuple[abi.Setup, abi.Batch]:
    """Dense batch equivalent to one bundled fixture: Y lists = dump "floor1 ys", spectra = dump "after_residue",
    emit counts = lengths of the dump's "pcm" entries. Coupling: stereo fixture has one step (0,1)."""
    C = int(g["channels"])
    bs = sorted(set(int(x) for x in g["blocksize"]))
    bs0, bs1 = bs[0], bs[-1]
    floors = [abi.Floor1(list(g["floor_xs"][i, :g["floor_nposts"][i]]), int(g["floor_multipliers"][i]))
              for i in range(len(g["floor_nposts"]))]
    coupl = [(0, 1)] if C == 2 else []
    mappings = [abi.Mapping([0] * C, [0], [0], coupl), abi.Mapping([0] * C, [1], [0], coupl)]
    setup = abi.Setup(C, int(g["sample_rate"]), (bs0, bs1), floors, mappings, [abi.Mode(0, 0), abi.Mode(1, 1)])
    n = g["blocksize"].astype(np.int64)
    P = len(n)
    blockflag = (n == bs1).astype(np.uint8)
    plan = plan_stream(blockflag, (bs0, bs1))
    emit = g["emit_frames"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(emit)[:-1]])
    packets = np.zeros(P, abi.PACKET_DTYPE)
    packets["mode"] = blockflag
    packets["window_flags"] = plan.window_flags
    packets["emit_frames"] = emit
    packets["pcm_off"] = off
    used = g["floor_used"]
    packets["floor_used"] = (used.astype(np.uint16) << np.arange(C, dtype=np.uint16)).sum(1)
    ys_list = []
    yo = 0
    for p in range(P):
        packets["ys_off"][p] = yo
        for c in range(C):
            if used[p, c]:
                k = int(g["floor_nposts"][g["floor_number"][p, c]])
                ys_list.append(g["ys"][p, c, :k].astype(np.uint16))
                yo += k
    packets["spec_off"] = np.concatenate([[0], np.cumsum(n // 2 * C)[:-1]])
    streams = np.zeros(1, abi.STREAM_DTYPE)
    streams[0] = (0, 0, P, 0, int(emit.sum()), 0)
    batch = abi.Batch(streams, packets, np.concatenate(ys_list), g["after_residue"].astype(np.float32),
                      int(emit.sum()) * C)
    return setup, batch


# ---- randomised setups (parity stress: many floors, submaps, modes, coupling graphs) ------------------------------
def random_setup(rng: np.random.Generator, channels: int, blocksizes=(256, 2048), max_posts: int = 32,
                 max_couplings: int = 5, max_component: int = 4, max_posts_short: Optional[int] = None) -> abi.Setup:
    """A random but well-formed setup: 1-2 floors per blocksize class with random distinct X lists and multipliers,
    1-3 submaps with their own floors, up to `max_couplings` coupling steps whose connected components stay within
    `max_component` channels, and 2-4 modes over 2-4 mappings. Floor class of a mapping follows its modes."""
    bs0, bs1 = blocksizes
    floors, cls_floors = [], {0: [], 1: []}
    for cls, half in ((0, bs0 // 2), (1, bs1 // 2)):
        for _ in range(int(rng.integers(1, 3))):
            cap = max_posts if (cls == 1 or max_posts_short is None) else max_posts_short
            posts = int(rng.integers(2, min(cap, half) + 1))
            inner = rng.choice(np.arange(1, half), size=posts - 2, replace=False) if posts > 2 else np.zeros(0, int)
            xs = [0, half] + [int(v) for v in inner]
            cls_floors[cls].append(len(floors))
            floors.append(abi.Floor1(xs, int(rng.integers(1, 5))))
    # coupling steps with bounded connected components
    comp = list(range(channels))
    size = {c: 1 for c in range(channels)}

    def find(c):
        while comp[c] != c:
            c = comp[c]
        return c
    coupl = []
    for _ in range(int(rng.integers(0, max_couplings + 1)) if channels > 1 else 0):
        m, a = (int(v) for v in rng.choice(channels, size=2, replace=False))
        rm, ra = find(m), find(a)
        if rm != ra:
            if size[rm] + size[ra] > max_component:
                continue
            comp[ra] = rm
            size[rm] += size[ra]
        coupl.append((m, a))
    n_submaps = int(rng.integers(1, min(3, channels) + 1))
    mux = [int(v) for v in rng.integers(0, n_submaps, size=channels)]
    mappings, modes = [], []
    n_modes = int(rng.integers(2, 5))
    flags = [0, 1] + [int(v) for v in rng.integers(0, 2, size=n_modes - 2)]
    for f in flags:
        mappings.append(abi.Mapping(mux=mux, submap_floor=[int(rng.choice(cls_floors[f])) for _ in range(n_submaps)],
                                    submap_residue=[0] * n_submaps, couplings=list(coupl)))
        modes.append(abi.Mode(f, len(mappings) - 1))
    return abi.Setup(channels=channels, sample_rate=44100, blocksize=(bs0, bs1), floors=floors, mappings=mappings, modes=modes)


def random_batch(setup: abi.Setup, rng: np.random.Generator, streams: int = 3, packets_per_stream: int = 70,
                 p_unused: float = 0.15, p_short: float = 0.2) -> abi.Batch:
    """Dense batch for an arbitrary setup (per-channel floors through the submaps, any mode of the right block class)."""
    C = setup.channels
    modes_of = {0: [i for i, m in enumerate(setup.modes) if not m.blockflag], 1: [i for i, m in enumerate(setup.modes) if m.blockflag]}
    pk_rows, st_rows, ys_all, spec_all = [], [], [], []
    ys_off = spec_off = pcm_base = first = 0
    for s in range(streams):
        bf = block_sequence(packets_per_stream, rng, p_short=p_short)
        plan = plan_stream(bf, setup.blocksize, trim_last=int(rng.integers(0, 40)))
        for k in range(packets_per_stream):
            mode = int(rng.choice(modes_of[int(bf[k])]))
            mp = setup.mappings[setup.modes[mode].mapping]
            half = int(plan.n[k]) // 2
            used = 0
            this_ys_off = ys_off
            draw = rng.random(C) >= p_unused
            if not closed_under_propagation(draw[None, :], mp.couplings)[0]:
                draw[:] = True
            for c in range(C):
                if draw[c]:
                    used |= 1 << c
                    fl = setup.floors[mp.submap_floor[mp.mux[c]]]
                    ys_all.append(gen_ys(rng, 1, fl.xs, fl.multiplier)[0])
                    ys_off += len(fl.xs)
            ub = np.array([[(used >> c) & 1 for c in range(C)]], bool)
            spec_all.append((gen_spectra(rng, C, half) * np.where(propagated(ub, mp.couplings)[0], np.float32(1), QUIET)[:, None]).ravel())
            pk_rows.append((s, mode, int(plan.window_flags[k]), used, int(plan.emit[k]), 0, int(plan.pcm_off[k]), this_ys_off, spec_off))
            spec_off += C * half
        st_rows.append((0, first, packets_per_stream, 0, plan.frames, pcm_base))
        pcm_base += plan.frames * C
        first += packets_per_stream
    packets = np.array(pk_rows, dtype=abi.PACKET_DTYPE)
    st = np.array(st_rows, dtype=abi.STREAM_DTYPE)
    ys = np.concatenate(ys_all).astype(np.uint16) if ys_all else np.zeros(0, np.uint16)
    return abi.Batch(streams=st, packets=packets, ys=ys, payload=np.concatenate(spec_all).astype(np.float32),
                     pcm_floats=pcm_base, input_kind=abi.POV_INPUT_DENSE, pcm_layout=abi.POV_PCM_PLANAR)


def concat_batches(a: abi.Batch, b: abi.Batch, setup_id_b: int) -> abi.Batch:
    """Streams of `b` appended to `a` (dense batches): arenas concatenated, offsets rebased, b's streams get setup_id_b."""
    assert a.input_kind == b.input_kind == abi.POV_INPUT_DENSE and a.pcm_layout == b.pcm_layout
    sb, pb = b.streams.copy(), b.packets.copy()
    sb["setup_id"] = setup_id_b
    sb["first_packet"] += len(a.packets)
    sb["pcm_base"] += np.uint64(a.pcm_floats)
    pb["stream"] += len(a.streams)
    pb["ys_off"] += np.uint64(len(a.ys))
    pad = (-len(a.payload)) % 4                       # keep 16-byte alignment of every spectrum
    pb["spec_off"] += np.uint64(len(a.payload) + pad)
    payload = np.concatenate([a.payload, np.zeros(pad, np.float32), b.payload])
    return abi.Batch(streams=np.concatenate([a.streams, sb]), packets=np.concatenate([a.packets, pb]),
                     ys=np.concatenate([a.ys, b.ys]), payload=payload, pcm_floats=a.pcm_floats + b.pcm_floats,
                     input_kind=a.input_kind, pcm_layout=a.pcm_layout)
