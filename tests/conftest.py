import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# written by tools/vorbis_writer.py and decoded by the unmodified reference (tests/golden/make_synthetic_golden.py);
# "shared_submap" is decoded by the hook-patched libvorbis instead (the reference is off-spec there, hpp:755)
SYNTHETIC = ("synth_res0_mono", "synth_two_submaps", "synth_surround51", "synth_codebooks", "synth_shared_submap",
             "synth_residue_edges")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    out = {}
    for name in ("stereo44khz", "mono44khz") + SYNTHETIC:
        with np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")) as z:
            out[name] = {k: z[k] for k in z.files}
    return out
