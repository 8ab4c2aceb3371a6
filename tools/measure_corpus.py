#!/usr/bin/env python3
"""Config 5 of BASELINE.json: sharded corpus decode — the bundled stereo fixture replicated N times, real host front
end (Ogg framing + Huffman/codebook decode, multi-threaded) -> descriptor batches -> GPU -> PCM back in host memory.
Prints end-to-end PCM samples/s next to the reference decoder on the same host cores. One GPU per process; with
torchrun every rank decodes its own shard of the file list (sharding.shard_range), no collective on the data path.

    python tools/measure_corpus.py [--files 10000] [--threads 0]"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parseoggvorbis_b200 import sharding  # noqa: E402
from parseoggvorbis_b200.lib import SynthContext  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=10000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--no-reference", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ogg = open(os.path.join(ROOT, "tests", "golden", "test.stereo44khz.ogg"), "rb").read()
    lo, hi = sharding.shard_range(a.files, world, rank)
    files = [ogg] * (hi - lo)
    # host threads of this rank: its share of the node's cores, minus one for the thread that feeds the GPU
    threads = a.threads or max(1, (os.cpu_count() or 2) // world - 1)
    ctx = SynthContext(local)
    # warm-up: tables, device arenas and the pinned staging pool of a long-lived decoder (pov_decode_corpus keeps them on the
    # context); 2048 files = 32 chunks is enough for every worker to have allocated its staging buffers
    ctx.decode_corpus(files[:min(2048, len(files))], threads)
    t0 = time.perf_counter()
    frames, total, chk = ctx.decode_corpus(files, threads)
    dt = time.perf_counter() - t0
    out = {"config": "config5: stereo fixture x %d (rank %d/%d decodes %d files)" % (a.files, rank, world, hi - lo),
           "host_threads": threads, "samples": total, "seconds": dt, "samples_per_s": total / dt,
           "frames_per_file": int(frames[0]) if len(frames) else 0, "checksum": chk}
    if rank == 0 and not a.no_reference:
        exe = os.path.join(ROOT, "oracle", "_ref", "ref_decode_bench")
        if os.path.exists(exe):
            cores = os.cpu_count() or 1
            r = json.loads(subprocess.check_output([exe, os.path.join(ROOT, "tests", "golden", "test.stereo44khz.ogg"), str(cores), "40"],
                                                   text=True).strip().splitlines()[-1])
            out["reference_cpu"] = {"cores": cores, "samples_per_s": r["samples"] / r["seconds"]}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
