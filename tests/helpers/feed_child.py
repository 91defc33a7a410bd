"""feed_child.py MODE WORKDIR -- child process of tests/test_gpu_dropin.py::test_feed_cufile.

Decodes a small clip with Decoder::loadFramesToDevice under MCRAW_FEED=MODE, checks every frame against its source image and
prints one JSON line {"feed": ..., "frames_ok": ...}.  It runs as a child because a feed that calls into a third-party
library (libcufile) may leave a helper thread behind that never returns: the parent sets the deadline, the child leaves
with os._exit so that no finaliser waits for that thread."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

mode, workdir = sys.argv[1], sys.argv[2]
os.environ["MCRAW_FEED"] = mode
os.environ["MCRAW_FEED_CHUNK_KB"] = "16"
os.environ.setdefault("MCRAW_CUFILE_DEADLINE_S", "8")
from motioncam_decoder_b200 import capi, hostapi, testvec as tv  # noqa: E402

frames, images = [], {}
for k in range(6):
    w, h = [(1928, 16), (640, 12), (4080, 8)][k % 3]
    img = tv.gen_photon(w, h, 1023, seed=70 + k)
    ts = 1000 + 33 * k
    legacy = k % 3 == 1
    frames.append({"timestamp": ts, "data": tv.encode_legacy(img) if legacy else tv.encode_current(img), "width": w, "height": h,
                   "compressionType": 6 if legacy else 7})
    images[ts] = img
path = os.path.join(workdir, "feed_child.mcraw")
tv.write_mcraw(path, frames)
dec = hostapi.Decoder(path)
stamps = dec.get_frames()
ctx = capi.Context(0)
ptrs = [ctx.device_alloc(images[ts].size * 2) for ts in stamps]
caps = [images[ts].size for ts in stamps]
ok = True
for _ in range(2):
    for p, ts in zip(ptrs, stamps):
        ctx.h2d(p, np.zeros(images[ts].shape, np.uint16))
    dec.load_frames_to_device(stamps, ptrs, caps)
    for ts, p in zip(stamps, ptrs):
        out = np.empty(images[ts].shape, dtype=np.uint16)
        ctx.d2h(out, p)
        ok = ok and bool(np.array_equal(out, images[ts]))
print(json.dumps({"feed": dec.feed_description(), "frames_ok": ok}), flush=True)
os._exit(0)
