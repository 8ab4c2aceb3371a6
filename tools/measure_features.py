#!/usr/bin/env python3
"""Feature-matrix throughput on BASELINE.json configs[3] (16 kHz mono speech clips, 10^6 packets: the returnn_import use
case): staged kernels + pov_batch_features for every stream of the batch, device-resident input, only the matrices cross
PCIe. Context for DESIGN.md, not a bench line.  python tools/measure_features.py > gpurun_out/features.json"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from parseoggvorbis_b200 import workloads  # noqa: E402
from parseoggvorbis_b200.lib import SynthContext  # noqa: E402


def main():
    ctx = SynthContext(0)
    setup, batch = workloads.config4(clips=10000, packets_per_clip=100, blocksizes=(256, 2048))
    batch.streams["setup_id"] = ctx.register_setup(setup)
    bh = ctx.upload(batch)
    ctx.sync(bh)
    out = {"config": "config4: 10000 mono clips x 100 packets, 256/2048", "packets": int(len(batch.packets))}
    t0 = time.perf_counter()
    ctx.run_staged(bh)
    ctx.sync(bh)
    out["staged_kernels_s"] = time.perf_counter() - t0
    ALL = 0xFFFFFFFF
    for kind, k in SynthContext.FEATURE_KINDS.items():
        dim = 64
        rows = C.c_uint64(0)
        ctx._check(ctx.L.pov_batch_features(ctx.ctx, bh.h, ALL, k, dim, None, 0, C.byref(rows)))
        m = np.zeros((int(rows.value), dim), np.float32)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ctx._check(ctx.L.pov_batch_features(ctx.ctx, bh.h, ALL, k, dim, m.ctypes.data_as(C.POINTER(C.c_float)), rows.value, C.byref(rows)))
            best = min(best, time.perf_counter() - t0)
        out[kind] = {"rows": int(rows.value), "dim": dim, "seconds": best, "rows_per_s": int(rows.value) / best,
                     "matrix_MB": m.nbytes / 1e6, "nonzero": int(np.count_nonzero(m))}
    print(json.dumps(out))
    bh.free()
    ctx.close()


if __name__ == "__main__":
    main()
