#!/usr/bin/env python3
"""SURVEY §8f-3 through the dump seam: the reference's own feature readers (demo_live_extract.CallbacksOutputReader
.read_floor_ys / .read_residue_ys, the functions returnn_import.get_features_from_raw_bytes ends in) run UNCHANGED on a
debug dump written by pov_ogg_vorbis_decode_memory on the B200 and must give the same matrices as on the reference
decoder's dump of the same file.

Runs only where /root/reference exists (this container), on dumps the GPU tests bring back in gpurun_out/:
    python tools/check_features_with_reference_readers.py [stereo44khz|mono44khz ...] > profiles/rNN_features_via_reference_readers.log
"""
import importlib
import os
import subprocess
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    names = sys.argv[1:] or ["stereo44khz", "mono44khz"]
    # better_exchook is not in this image; the reference uses it for tracebacks only
    sys.modules.setdefault("better_exchook", types.SimpleNamespace(install=lambda *a, **k: None, better_exchook=None))
    sys.path.insert(0, os.path.dirname("/root/reference"))
    dle = importlib.import_module("reference.demo_live_extract")
    ok = True
    for name in names:
        ogg = os.path.join(ROOT, "tests", "golden", "test.%s.ogg" % name)
        ours = os.path.join(ROOT, "gpurun_out", "%s_b200.dbg" % name)
        with tempfile.TemporaryDirectory() as td:
            ref = os.path.join(td, "ref.dbg")
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ours.bin"), "--in", ogg, "--debug_out", ref], stdout=subprocess.DEVNULL)
            for label, call in (("read_floor_ys(output_dim=32)", lambda r: r.read_floor_ys(output_dim=32)),
                                ("read_floor_ys(output_dim=64)", lambda r: r.read_floor_ys(output_dim=64)),
                                ("read_residue_ys(output_dim=64)", lambda r: r.read_residue_ys(output_dim=64)),
                                ("read_residue_ys(output_dim=40, log1p_abs_space=True)", lambda r: r.read_residue_ys(output_dim=40, log1p_abs_space=True))):
                a = np.asarray(call(dle.CallbacksOutputReader(open(ref, "rb"))))
                b = np.asarray(call(dle.CallbacksOutputReader(open(ours, "rb"))))
                same = a.shape == b.shape and bool(np.array_equal(a, b))
                ok &= same
                print("%-12s %-55s shape %-12s %s" % (name, label, a.shape, "identical" if same else "DIFFERENT (max abs %g)" % float(np.abs(a - b).max())))
    print("OK" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
