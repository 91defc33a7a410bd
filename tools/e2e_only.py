"""e2e_only.py -- micro-benchmark of the host-input pipeline alone (pinned ring -> staged H2D -> decode), C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from motioncam_decoder_b200 import capi
desc, w, h, ct, _gen, _maxval, frames, _distinct, _strong = bench.WORKLOADS['c2']
streams, _ = bench.make_streams('c2', want_images=False)
ctx = capi.Context(0)
offs, total = [], 0
for i in range(frames):
    offs.append(total); total += (len(streams[i % len(streams)]) + 255) & ~255
ring_ptr, ring = ctx.pinned_array(total + 256)
items = []
pitch = (w * h * 2 + 511) & ~255
dev = ctx.device_alloc(frames * pitch + 256)                    # one output area, frames at a constant pitch (as in bench.py)
for i in range(frames):
    s = streams[i % len(streams)]
    ring[offs[i]:offs[i] + len(s)] = s
    items.append((ring_ptr + offs[i], len(s), w, h, ct, dev + i * pitch, w * h))
descs, n = capi.Context.make_descs(items)
HOST_OUT = len(sys.argv) > 1 and sys.argv[1] == "host-out"      # pixels back into pinned host memory (mcraw_decode_batch_host_out)
if HOST_OUT:
    import ctypes
    out_ptr, _ = ctx.pinned_array(frames * pitch + 256)
    host_out = (ctypes.c_void_p * frames)(*[out_ptr + i * pitch for i in range(frames)])
    step = lambda: ctx.decode_batch_host_out(descs, host_out, n)
else:
    step = lambda: ctx.decode_batch_host(descs, n)
for _ in range(3):
    step(); ctx.batch_wait(n)
t0 = time.perf_counter()
R = 30
for _ in range(R):
    step(); ctx.batch_wait(n)
t = (time.perf_counter() - t0) / R
comp = sum(len(streams[i % len(streams)]) for i in range(frames))
print("host-out" if HOST_OUT else "device-out", os.environ.get('MCRAW_HOSTOUT_FIRST_MB'), os.environ.get('MCRAW_STAGE_MB'), os.environ.get('MCRAW_STAGE_N'), os.environ.get('MCRAW_COPY_N'), f"{t*1e3:.3f} ms/step  {frames*w*h/t/1e9:.1f} Gpix/s  {comp/t/1e9:.1f} GB/s")
