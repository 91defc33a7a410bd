"""e2e_only.py -- micro-benchmark of the host-input pipeline alone (pinned ring -> staged H2D -> decode), C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motioncam_decoder_b200 import capi
desc, w, h, ct, frames, streams = bench.make_streams('c2')
ctx = capi.Context(0)
offs, total = [], 0
for i in range(frames):
    offs.append(total); total += (len(streams[i % len(streams)]) + 255) & ~255
ring_ptr, ring = ctx.pinned_array(total + 256)
items = []
for i in range(frames):
    s = streams[i % len(streams)]
    ring[offs[i]:offs[i] + len(s)] = s
    dp = ctx.device_alloc(w * h * 2 + 256)
    items.append((ring_ptr + offs[i], len(s), w, h, ct, dp, w * h))
descs, n = capi.Context.make_descs(items)
for _ in range(3):
    ctx.decode_batch_host(descs, n); ctx.batch_wait(n)
t0 = time.perf_counter()
R = 30
for _ in range(R):
    ctx.decode_batch_host(descs, n); ctx.batch_wait(n)
t = (time.perf_counter() - t0) / R
comp = sum(len(streams[i % len(streams)]) for i in range(frames))
print(os.environ.get('MCRAW_STAGE_MB'), os.environ.get('MCRAW_STAGE_N'), os.environ.get('MCRAW_COPY_N'), f"{t*1e3:.3f} ms/step  {frames*w*h/t/1e9:.1f} Gpix/s  {comp/t/1e9:.1f} GB/s")
