import sys, numpy as np
sys.path.insert(0, '/root/repo')
from parseoggvorbis_b200 import workloads
from parseoggvorbis_b200.lib import SynthContext
ctx = SynthContext(0)
for mk in (lambda: workloads.config2(P=300, streams=3, distinct=3, seed=0),
           lambda: workloads.config3(P=64, streams=2, distinct=2, seed=1),
           lambda: workloads.config4(clips=6, packets_per_clip=40)):
    setup, batch = mk()
    batch.streams["setup_id"] = ctx.register_setup(setup)
    bh = ctx.upload(batch); ctx.run(bh); pcm = ctx.fetch_pcm(bh)
    print(ctx.kernel_name(bh), pcm.size, float(np.abs(pcm).max()), int(ctx.status(bh).any()))
    bh.free()
rng = np.random.default_rng(1003)
s = workloads.random_setup(rng, 5); b = workloads.random_batch(s, rng, streams=2, packets_per_stream=40)
b.streams["setup_id"] = ctx.register_setup(s)
bh = ctx.upload(b); ctx.run(bh); print(ctx.kernel_name(bh), ctx.fetch_pcm(bh).size); bh.free()
ctx.close()
