#!/usr/bin/env python3
"""Regenerate the golden vectors under tests/golden/ from the UNMODIFIED reference decoder.

Run in the build container (needs /root/reference and a built oracle/_ref/, i.e. `make -C oracle ref`):

    python tests/golden/make_golden.py

For both bundled fixtures (/root/reference/tests/audio/*.ogg) it runs ``oracle/_ref/ours.bin --in F --debug_out D``
(the reference's own CLI, src/main.cpp:53-67) and stores what the dump holds per audio packet, regrouped into
flat arrays, as ``<name>.npz``. It also copies the two .ogg *data* fixtures (test inputs, not source code) so
that the GPU box — which has no /root/reference — can decode the same bytes, and extracts the 256-entry
floor1_inverse_dB_table the reference compiles in (src/inverse_db_table.h:13-78) by compiling a 5-line
program against that header.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.dumpfile import load_dump  # noqa: E402

REF = os.environ.get("POV_REFERENCE", "/root/reference")
OURS = os.path.join(ROOT, "oracle", "_ref", "ours.bin")
FIXTURES = {"stereo44khz": "test.stereo44khz.ogg", "mono44khz": "test.mono44khz.ogg"}


def pack(dump):
    C = dump.num_channels
    P = len(dump.packets)
    maxposts = max(len(x) for x in dump.floor_xs)
    out = {
        "sample_rate": np.uint32(dump.sample_rate),
        "channels": np.uint32(C),
        "floor_multipliers": np.array(dump.floor_multipliers, np.uint8),
        "floor_xs": np.zeros((len(dump.floor_xs), maxposts), np.uint32),
        "floor_nposts": np.array([len(x) for x in dump.floor_xs], np.uint32),
        "blocksize": np.array([p.blocksize for p in dump.packets], np.uint32),
        "abs_total_pos": np.array([p.abs_total_pos for p in dump.packets], np.uint64),
        "expected_ending_total_pos": np.array([p.expected_ending_total_pos for p in dump.packets], np.int64),
        "floor_number": np.zeros((P, C), np.uint8),
        "floor_used": np.zeros((P, C), bool),
        "ys": np.zeros((P, C, maxposts), np.uint32),
        "final_ys": np.zeros((P, C, maxposts), np.uint32),
        "step2_flag": np.zeros((P, C, maxposts), bool),
        "emit_frames": np.zeros(P, np.uint32),
        "has_floor_stages": np.bool_(True),
    }
    for i, x in enumerate(dump.floor_xs):
        out["floor_xs"][i, :len(x)] = x
    # variable-length float stages: concatenated per (packet, channel) in order
    floors, ares, aenv, mdct, pcm = [], [], [], [], []
    for pi, p in enumerate(dump.packets):
        for c in range(C):
            f = p.floors[c]
            out["floor_number"][pi, c] = f.floor_number
            if f.ys is not None:
                out["floor_used"][pi, c] = True
                n = len(f.ys)
                out["ys"][pi, c, :n] = f.ys
                if f.final_ys is None:       # the libvorbis hooks dump the coded Ys only (compare-debug-out.py:192-195)
                    out["has_floor_stages"] = np.bool_(False)
                    floors.append(np.zeros(p.blocksize, np.uint8))
                else:
                    out["final_ys"][pi, c, :n] = f.final_ys
                    out["step2_flag"][pi, c, :n] = f.step2_flag
                    assert f.floor.max() < 256
                    floors.append(f.floor.astype(np.uint8))
                    # floor_outputs is inverse_db_table[floor] (hpp:588): verified here, not stored
            else:
                floors.append(np.zeros(p.blocksize, np.uint8))
            ares.append(p.after_residue[c])
            aenv.append(p.after_envelope[c])
            mdct.append(p.pcm_after_mdct[c])
        out["emit_frames"][pi] = len(p.pcm[0]) if 0 in p.pcm else 0
    out["floor"] = np.concatenate(floors)
    out["after_residue"] = np.concatenate(ares)
    out["after_envelope"] = np.concatenate(aenv)
    out["pcm_after_mdct"] = np.concatenate(mdct)
    out["pcm"] = dump.pcm_concat()
    return out


def extract_inverse_db_table(tmp):
    src = os.path.join(tmp, "t.cpp")
    with open(src, "w") as f:
        f.write('#include <cstdio>\n#include "inverse_db_table.h"\n'
                'int main(){fwrite(inverse_db_table,4,256,stdout);'
                'return sizeof(inverse_db_table)==1024?0:1;}\n')
    exe = os.path.join(tmp, "t.bin")
    subprocess.check_call(["g++", "-I", os.path.join(REF, "src"), src, "-o", exe])
    raw = subprocess.check_output([exe])
    return np.frombuffer(raw, dtype="<f4").copy()


def main():
    assert os.path.exists(OURS), "run `make -C oracle ref` first"
    with tempfile.TemporaryDirectory() as tmp:
        table = extract_inverse_db_table(tmp)
        np.save(os.path.join(HERE, "inverse_db_table.npy"), table)
        for name, fn in FIXTURES.items():
            ogg = os.path.join(REF, "tests", "audio", fn)
            shutil.copyfile(ogg, os.path.join(HERE, fn))
            dbg = os.path.join(tmp, name + ".dbg")
            subprocess.check_call([OURS, "--in", ogg, "--debug_out", dbg], stdout=subprocess.DEVNULL)
            dump = load_dump(dbg)
            for p in dump.packets:  # floor_outputs == table[floor] on every rendered channel
                for f in p.floors.values():
                    if f.floor is not None:
                        assert np.array_equal(table[f.floor].view(np.uint32), f.floor_outputs.view(np.uint32))
            arrs = pack(dump)
            path = os.path.join(HERE, name + ".npz")
            np.savez_compressed(path, **arrs)
            print(name, "packets", len(dump.packets), "frames", arrs["pcm"].shape, "->",
                  os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
