#!/usr/bin/env python
"""export_e2e.py -- whole-program "dump a clip to DNG" (the reference's example.cpp workflow):
a synthetic .mcraw in tmpfs -> audio.wav + frame_%06d.dng in tmpfs, this repo's mcraw_export (batched B200 decode,
threaded writers) next to the compiled reference program (oracle/_ref/ref_example, CPU decode, one thread) when present.
Wall-clock of the whole process, CUDA context creation included; files compared byte for byte afterwards."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--ref-frames", type=int, default=8)
    ap.add_argument("--width", type=int, default=4080)
    ap.add_argument("--height", type=int, default=3072)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--dir", default="/dev/shm/mcraw_export_e2e")
    a = ap.parse_args()
    from motioncam_decoder_b200 import _lib, hostapi, testvec as tv
    import oracle_lib as ol
    shutil.rmtree(a.dir, ignore_errors=True)
    os.makedirs(a.dir)
    distinct = [tv.encode_current(tv.gen_photon(a.width, a.height, 1023, seed=1 + i)) for i in range(4)]
    frames = [{"timestamp": 1000 + i, "data": distinct[i % 4], "width": a.width, "height": a.height, "compressionType": 7}
              for i in range(a.frames)]
    rng = np.random.default_rng(1)
    audio = [(100 * k, rng.integers(-8000, 8000, 3840, dtype=np.int16)) for k in range(8)]
    clip = os.path.join(a.dir, "clip.mcraw")
    tv.write_mcraw(clip, frames, audio)
    out = {"frames": a.frames, "frame": f"{a.width}x{a.height}", "clip_bytes": os.path.getsize(clip), "batch": a.batch, "threads": a.threads}

    ours = os.path.join(a.dir, "ours")
    os.makedirs(ours)
    exe = os.path.join(_lib.PKG_DIR, "mcraw_export")
    t0 = time.perf_counter()
    r = subprocess.run([exe, clip, "--batch", str(a.batch), "--threads", str(a.threads)], cwd=ours, capture_output=True, text=True)
    t = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(r.stderr)
    out["mcraw_export"] = {"seconds": round(t, 3), "frames_per_s": round(a.frames / t, 1),
                           "mpix_per_s": round(a.frames * a.width * a.height / t / 1e6, 1)}
    # the same clip again inside this process (context creation and first-touch of the pinned buffers not repeated by the OS cache)
    again = os.path.join(a.dir, "again")
    os.makedirs(again)
    t0 = time.perf_counter()
    st = {}
    hostapi.export_clip(clip, again, batch=a.batch, writer_threads=a.threads, stats=st)
    t = time.perf_counter() - t0
    out["export_clip_in_process"] = {"seconds": round(t, 3), "frames_per_s": round(a.frames / t, 1),
                                     "mpix_per_s": round(a.frames * a.width * a.height / t / 1e6, 1), "stats": st,
                                     "steady_frames_per_s": round(a.frames / st["steady_s"], 1) if st.get("steady_s") else None}
    shutil.rmtree(again)

    ref_exe = os.path.join(os.path.dirname(ol.REF_SO), "ref_example")
    if os.path.exists(ref_exe):
        ref = os.path.join(a.dir, "ref")
        os.makedirs(ref)
        n = min(a.ref_frames, a.frames)
        t0 = time.perf_counter()
        r = subprocess.run([ref_exe, clip, "-n", str(n)], cwd=ref, capture_output=True, text=True)
        t = time.perf_counter() - t0
        if r.returncode != 0:
            raise SystemExit(r.stderr)
        out["reference_example"] = {"frames": n, "seconds": round(t, 3), "frames_per_s": round(n / t, 1),
                                    "mpix_per_s": round(n * a.width * a.height / t / 1e6, 1)}
        same = all(open(os.path.join(ref, f), "rb").read() == open(os.path.join(ours, f), "rb").read() for f in sorted(os.listdir(ref)))
        out["files_identical"] = bool(same)
    shutil.rmtree(a.dir, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
