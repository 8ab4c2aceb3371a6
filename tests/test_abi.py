"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol include/pov_synth.h declares,
the ctypes mirrors have the C compiler's struct sizes, and host-only helpers agree with the goldens."""
import ctypes as C
import os
import re
import subprocess

import numpy as np

from parseoggvorbis_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pov_synth.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pov_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    declared = _declared_symbols()
    assert sorted(lib.SYMBOLS) == declared
    L = lib.load()
    for sym in declared:
        assert getattr(L, sym) is not None
    assert L.pov_abi_version() == abi.POV_ABI_VERSION


def test_struct_sizes_match_c_compiler(tmp_path):
    names = ["pov_codebook", "pov_floor1", "pov_floor1_syntax", "pov_residue", "pov_mapping", "pov_mode", "pov_setup", "pov_stream",
             "pov_packet", "pov_batch", "pov_decoded"]
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "pov_synth.h"\nint main(){' +
                   "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names) + "return 0;}\n")
    exe = tmp_path / "s.bin"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(getattr(abi, n)), n


def test_inverse_db_table_matches_reference_golden():
    t = np.zeros(256, np.float32)
    lib.load().pov_inverse_db_table(t.ctypes.data_as(C.POINTER(C.c_float)))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "inverse_db_table.npy"))
    assert np.array_equal(t.view(np.uint32), gold.view(np.uint32))


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        return
    ctx = C.c_void_p(None)
    err = C.c_char_p(None)
    rc = lib.load().pov_ctx_create(0, C.byref(ctx), C.byref(err))
    assert rc == abi.POV_ERR_CUDA and not ctx.value
    assert b"no CPU fallback" in err.value or b"CUDA" in err.value


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under parseoggvorbis_b200/ may import, link or load it."""
    pkg = os.path.join(ROOT, "parseoggvorbis_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or fn == "Makefile":
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in txt, os.path.join(dirpath, fn)
    out = subprocess.check_output(["ldd", lib.LIB_PATH], text=True)
    assert "oracle" not in out


def test_windows_bit_exact_for_every_block_size_pair_and_flag_combination():
    """a8 (hpp:837-862): the window the overlap-add kernels multiply by — built from the library's slope tables — equals
    the restated reference expression bit for bit: all block-size pairs 64..8192, short and long, all prev/next flags."""
    from tests import oracle_binding as ob
    L = lib.load()
    sizes = [64, 128, 256, 512, 1024, 2048, 4096, 8192]
    checked = 0
    for bs0 in sizes:
        for bs1 in sizes:
            if bs0 > bs1:
                continue
            for blockflag in (0, 1):
                n = bs1 if blockflag else bs0
                for prev in (0, 1):
                    for nxt in (0, 1):
                        got = np.zeros(n, np.float32)
                        rc = L.pov_window(bs0, bs1, blockflag, prev, nxt, got.ctypes.data_as(C.POINTER(C.c_float)), n)
                        assert rc == 0, (bs0, bs1, blockflag, prev, nxt)
                        ref = ob.window(bs0, bs1, blockflag, prev, nxt)
                        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (bs0, bs1, blockflag, prev, nxt)
                        checked += 1
    assert checked == 36 * 8
    bad = np.zeros(256, np.float32)
    assert L.pov_window(256, 128, 0, 0, 0, bad.ctypes.data_as(C.POINTER(C.c_float)), 256) != 0      # bs0 > bs1
    assert L.pov_window(96, 2048, 0, 0, 0, bad.ctypes.data_as(C.POINTER(C.c_float)), 96) != 0        # not a power of two
