#!/usr/bin/env python
"""Config 5 of BASELINE.json: a .mcraw FILE -> pinned host ring -> overlapped H2D -> decode -> 16-bit device buffers,
frame-parallel over the GPUs of one box, audio chunks passed through on the host.

    python tools/mcraw_file_e2e.py [--frames 64] [--workload c3|c2] [--reps 5] [--dir /dev/shm]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/mcraw_file_e2e.py ...

Rank 0 writes a synthetic clip (the test-vector encoder + container writer), every rank opens it with the drop-in
motioncam::Decoder, takes a contiguous shard of the timestamp-sorted frame list (shard.py) and decodes it with
Decoder::loadFramesToDevice.  Prints one JSON line (whole-job Mpix/s, file bytes/s, max over ranks).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--workload", default="c3", choices=["c2", "c3"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dir", default="/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["MCRAW_B200_DEVICE"] = str(local)       # device of the drop-in library's per-thread context

    sys.stdout.flush()
    json_fd = os.dup(1)            # libraries (NCCL) print to the C-level stdout: keep it for the one JSON line
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from motioncam_decoder_b200 import capi, hostapi, shard, testvec as tv
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    w, h, maxv, gen = (4080, 3072, 1023, "flatnoise") if args.workload == "c3" else (1920, 1080, 4095, "photon")
    path = os.path.join(args.dir, f"mcraw_e2e_{args.workload}_{args.frames}.mcraw")
    rng = np.random.default_rng(7)
    audio = [(1_000_000 * i if i % 2 == 0 else None, rng.integers(-3000, 3000, 1920 * 2, dtype=np.int16)) for i in range(16)]
    distinct = 4
    images = [tv.gen_flatnoise(w, h, 256, seed=s + 1) if gen == "flatnoise" else tv.gen_photon(w, h, maxv, seed=s + 1)
              for s in range(distinct)]
    if rank == 0:
        streams = [tv.encode_current(im) for im in images]
        frames = [{"timestamp": 1000 + 33 * i, "data": streams[i % distinct], "width": w, "height": h, "compressionType": 7}
                  for i in range(args.frames)]
        tv.write_mcraw(path, frames, audio)
    if world > 1:
        dist.barrier()

    dec = hostapi.Decoder(path)
    stamps = dec.get_frames()
    mine = [stamps[i] for i in shard.shard_contiguous(len(stamps), world, rank)]
    ctx = capi.Context(local)                          # device buffers for this rank's shard
    ptrs = [ctx.device_alloc(w * h * 2) for _ in mine]
    caps = [w * h] * len(mine)
    dec.load_frames_to_device(mine, ptrs, caps)         # warm-up (ring allocation, page cache)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        dec.load_frames_to_device(mine, ptrs, caps)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    # spot check: first and last frame of the shard against the source images
    ok = True
    out = np.empty((h, w), dtype=np.uint16)
    for k in (0, len(mine) - 1):
        if not mine:
            break
        ctx.d2h(out, ptrs[k])
        ok &= bool(np.array_equal(out, images[stamps.index(mine[k]) % distinct]))
    audio_ok = None
    if rank == 0:
        got = dec.load_audio()
        audio_ok = [(t, d.tobytes()) for t, d in got] == [(-1 if t is None else t, np.asarray(d).tobytes()) for t, d in audio]
    if rank == 0:
        size = os.path.getsize(path)
        os.write(json_fd, (json.dumps({"metric": "file_to_device_mpix_per_s", "value": args.frames * w * h * args.reps / dt / 1e6, "unit": "Mpix/s",
                          "n_gpus": world, "frames": args.frames, "reps": args.reps, "workload": args.workload,
                          "file_bytes": size, "file_gb_per_s": size * args.reps / dt / 1e9, "frames_ok": ok, "audio_ok": audio_ok,
                          "path": "file (page cache) -> pread into pinned ring -> staged H2D -> k_meta/k_units -> device u16"}) + "\n").encode())
        try:
            os.remove(path)
        except OSError:
            pass
    assert ok
    for p in ptrs:
        ctx.device_free(p)
    dec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
