import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from parseoggvorbis_b200 import workloads, abi
from parseoggvorbis_b200.lib import SynthContext
setup, batch = workloads.config2(P=4096, streams=32, distinct=4, seed=0)
ctx = SynthContext(0)
batch.streams["setup_id"] = ctx.register_setup(setup)
stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", 0))
for layout in (abi.POV_PCM_PLANAR, abi.POV_PCM_INTERLEAVED):
    b2 = abi.Batch(batch.streams.copy(), batch.packets, batch.ys, batch.payload, batch.pcm_floats, batch.input_kind, layout)
    if layout == abi.POV_PCM_INTERLEAVED:
        pass
    bh = ctx.upload(b2)
    for _ in range(3): ctx.run(bh)
    ctx.sync(bh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10): ctx.run(bh)
        e1.record(stream)
    ctx.sync(bh); torch.cuda.synchronize()
    print("layout", layout, ctx.kernel_name(bh), "ms", e0.elapsed_time(e1)/10, "Gsamples/s", batch.pcm_floats/(e0.elapsed_time(e1)/10*1e-3)/1e9)
    bh.free()
ctx.close()
