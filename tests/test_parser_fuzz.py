"""Memory-safety of the host front end on hostile input: tests/native/fuzz_parser.cpp under ASan + UBSan.
(compute-sanitizer is not available on the GPU pool; the parser is the component that reads untrusted bytes.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "parseoggvorbis_b200", "csrc")


@pytest.fixture(scope="module")
def fuzzer(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path_factory.mktemp("fuzz") / "fuzz_parser")
    cmd = ["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-std=c++17",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           os.path.join(ROOT, "tests", "native", "fuzz_parser.cpp"), os.path.join(CSRC, "vorbis_parse.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build unavailable: " + r.stderr[-300:])
    return exe


@pytest.mark.parametrize("name,iters,seed,mode", [
    ("stereo44khz", 1200, 1, None), ("mono44khz", 1200, 2, None),
    ("stereo44khz", 800, 3, "hdr"), ("mono44khz", 800, 4, "hdr"),
])
def test_parser_survives_mutated_files(fuzzer, name, iters, seed, mode):
    args = [fuzzer, os.path.join(ROOT, "tests", "golden", "test.%s.ogg" % name), str(iters), str(seed)] + ([mode] if mode else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert "fuzz:" in r.stdout and "runtime error" not in r.stderr
