#!/usr/bin/env python3
"""Device-resident kernel timings of the parity-test configurations of BASELINE.json (configs[2..4]) — not bench lines,
context for DESIGN.md. Usage (GPU box):  python tools/measure_configs.py > gpurun_out/configs.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parseoggvorbis_b200 import workloads  # noqa: E402
from parseoggvorbis_b200.lib import SynthContext  # noqa: E402


def timed(ctx, bh, steps=10, warm=3):
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", 0))
    for _ in range(warm):
        ctx.run(bh)
    ctx.sync(bh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            ctx.run(bh)
        e1.record(stream)
    ctx.sync(bh)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    torch.cuda.set_device(0)
    ctx = SynthContext(0)
    out = []
    cases = [
        ("config3: 48 kHz 5.1, long blocks, 3 coupling steps, 2 submaps", lambda: workloads.config3(P=4096, streams=64, distinct=2, seed=1)),
        ("config4a: 16 kHz mono clips 256/2048, 10000 clips x 100 packets", lambda: workloads.config4(clips=10000, packets_per_clip=100, blocksizes=(256, 2048))),
        ("config4b: 16 kHz mono clips 512/1024, 10000 clips x 100 packets", lambda: workloads.config4(clips=10000, packets_per_clip=100, blocksizes=(512, 1024))),
    ]
    for name, make in cases:
        t0 = time.time()
        setup, batch = make()
        batch.streams["setup_id"] = ctx.register_setup(setup)
        bh = ctx.upload(batch)
        ms = timed(ctx, bh)
        status_bad = int(np.count_nonzero(ctx.status(bh)))
        samples = int(batch.pcm_floats)
        out.append({"config": name, "kernel": ctx.kernel_name(bh), "packets": int(len(batch.packets)), "samples": samples,
                    "ms_per_launch": ms, "samples_per_s": samples / (ms * 1e-3), "GBps_8B_per_sample": samples * 8 / (ms * 1e-3) / 1e9,
                    "packets_with_status": status_bad, "gen_s": round(time.time() - t0, 1)})
        bh.free()
        print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
