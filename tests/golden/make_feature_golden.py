#!/usr/bin/env python3
"""Golden feature matrices for SURVEY.md §8(f)-3: what returnn_import.ParseOggVorbisLib.get_features_from_raw_bytes returns
(reference: returnn_import.py:74-115) for the four kinds this repo produces on the device, computed by the reference's OWN
readers (demo_live_extract.CallbacksOutputReader.read_floor_ys / read_residue_ys, demo_live_extract.py:262-505) on the
reference decoder's dump of each fixture, filtered to the entry names each kind asks for.

    python tests/golden/make_feature_golden.py        (build container only: imports /root/reference, needs oracle/_ref)
writes tests/golden/features.npz  (keys "<fixture>/<kind>/<output_dim>").
"""
import importlib
import os
import subprocess
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FIXTURES = {"stereo44khz": "test.stereo44khz.ogg", "mono44khz": "test.mono44khz.ogg", "synth_two_submaps": "synth_two_submaps.ogg",
            "synth_surround51": "synth_surround51.ogg"}
# kind -> (entry names the reference filters for, reader method)      (returnn_import.py:84-113)
KINDS = {
    "floor_final_ys": (["floor1_unpack multiplier", "floor1_unpack xs", "finish_setup", "floor_number", "floor1 final_ys",
                        "finish_audio_packet"], "read_floor_ys"),
    "floor_final_ys_rendered": (["floor1_unpack multiplier", "floor1_unpack xs", "finish_setup", "floor_number", "floor1 floor",
                                 "finish_audio_packet"], "read_floor_ys"),
    "residue_ys": (["floor1_unpack multiplier", "floor1_unpack xs", "finish_setup", "floor_number", "after_residue",
                    "finish_audio_packet"], "read_residue_ys"),
    "residue_ys_with_floor": (["floor1_unpack multiplier", "floor1_unpack xs", "finish_setup", "floor_number", "floor1 floor",
                               "after_residue", "finish_audio_packet"], "read_residue_ys"),
}
DIMS = (16, 40, 64)


def main():
    sys.modules.setdefault("better_exchook", types.SimpleNamespace(install=lambda *a, **k: None, better_exchook=None))
    sys.path.insert(0, os.path.dirname("/root/reference"))
    dle = importlib.import_module("reference.demo_live_extract")

    class Filtered(dle.CallbacksOutputReader):
        """The stream a decoder registered with set_data_filter(names) would have written (src/Callbacks.cpp:224-242)."""
        names = ()

        def read_entry(self):
            while True:
                name, channel, data = super().read_entry()
                if name in self.names:
                    return name, channel, data

    out = {}
    with tempfile.TemporaryDirectory() as td:
        for fx, fn in FIXTURES.items():
            dbg = os.path.join(td, fx + ".dbg")
            subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ours.bin"), "--in", os.path.join(HERE, fn), "--debug_out", dbg],
                                  stdout=subprocess.DEVNULL)
            for kind, (names, method) in KINDS.items():
                for dim in DIMS:
                    if method == "read_residue_ys" and dim < 32:
                        continue          # the reader asserts output_dim >= posts of the biggest floor (demo_live_extract.py:486)
                    r = Filtered(open(dbg, "rb"))
                    r.names = set(names)
                    m = np.asarray(getattr(r, method)(output_dim=dim), np.float32)
                    out["%s/%s/%d" % (fx, kind, dim)] = m
                    print("%-20s %-26s dim %3d -> %s" % (fx, kind, dim, m.shape))
    np.savez_compressed(os.path.join(HERE, "features.npz"), **out)
    print("wrote", os.path.join(HERE, "features.npz"), os.path.getsize(os.path.join(HERE, "features.npz")), "bytes")


if __name__ == "__main__":
    main()
