"""CPU checks of the synthetic workload generators against the oracle (no GPU): descriptors are well formed,
the oracle decodes them without tripping any reference CHECK, replicas are exact copies."""
import numpy as np

from parseoggvorbis_b200 import abi, workloads
from tests import oracle_binding as ob


def _run(setup, batch):
    pcm, status = ob.synth_batch([setup], batch, imdct="fast")
    assert not status.any()
    assert np.isfinite(pcm).all() and np.abs(pcm).max() > 0
    return pcm


def test_config2_generator_and_replication():
    setup, b = workloads.config2(P=200, streams=4, distinct=2, seed=0)
    assert len(b.streams) == 4 and len(b.packets) == 800
    assert (b.packets["spec_off"] % 4 == 0).all()
    pcm = _run(setup, b)
    per = b.pcm_floats // 2
    assert np.array_equal(pcm[:per], pcm[per:])       # replica of the two distinct streams
    n = np.asarray(setup.blocksize)[b.packets["mode"].astype(int)]
    assert set(np.unique(n)) == {256, 2048}


def test_config3_and_config4_generators():
    setup, b = workloads.config3(P=60, streams=2, distinct=2)
    assert setup.channels == 6 and len(setup.mappings[0].couplings) == 3
    _run(setup, b)
    for bs in [(256, 2048), (512, 1024), (64, 8192)]:
        setup, b = workloads.config4(clips=5, packets_per_clip=30, blocksizes=bs)
        _run(setup, b)


def test_golden_batch_roundtrip_shapes(golden):
    for g in golden.values():
        setup, b = workloads.golden_setup_and_batch(g)
        assert b.pcm_floats == g["pcm"].size
        assert len(b.payload) == g["after_residue"].size


def test_trimmed_last_packet_and_interleaved_layout():
    setup, _ = workloads.config2(P=8)
    rng = np.random.default_rng(3)
    plan = workloads.plan_stream(workloads.block_sequence(50, rng), setup.blocksize, trim_last=77)
    a = workloads.build_dense_batch(setup, [plan], np.random.default_rng(4))
    b = workloads.build_dense_batch(setup, [plan], np.random.default_rng(4), pcm_layout=abi.POV_PCM_INTERLEAVED)
    pa, pb = _run(setup, a), _run(setup, b)
    assert np.array_equal(pa.reshape(2, -1), pb.reshape(-1, 2).T)
