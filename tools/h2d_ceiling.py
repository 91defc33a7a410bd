#!/usr/bin/env python
"""h2d_ceiling.py -- what the host side of the box can feed: pinned host -> device copy bandwidth per GPU while ALL ranks
copy at once, nothing else running.  The ceiling against which bench.py's end-to-end legs are read (e2e.h2d_peak_gbs is
the same measurement taken inside the bench run).

    python tools/h2d_ceiling.py                                                            # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py

Per rank: `--streams` copy streams, each cycling over its own pinned 96 MB chunk (the staging size mcraw_decode_batch_host
uses) into its own device buffer; optional --d2h adds device -> host copies on further streams (full duplex).
Prints one JSON line: GB/s per rank (min / mean / max) and the aggregate, for H2D alone and for the duplex case.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk-mb", type=int, default=96)
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--streams", type=int, default=2)
    a = ap.parse_args()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from motioncam_decoder_b200 import numa
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    props = torch.cuda.get_device_properties(local)
    try:
        placement = numa.bind_to_gpu_node(f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0")
    except AttributeError:
        placement = "not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.chunk_mb << 20
    hosts = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(a.streams)]
    devs = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.streams)]
    hback = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(a.streams)]
    dback = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(a.streams)]
    up = [torch.cuda.Stream() for _ in range(a.streams)]
    down = [torch.cuda.Stream() for _ in range(a.streams)]

    def run(duplex):
        def go(reps):
            for r in range(reps):
                for k in range(a.streams):
                    with torch.cuda.stream(up[k]):
                        devs[k].copy_(hosts[k], non_blocking=True)
                    if duplex:
                        with torch.cuda.stream(down[k]):
                            hback[k].copy_(dback[k], non_blocking=True)
        go(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in up + down:
            s.wait_event(e0)
        go(a.reps)
        for s in up + down:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        return n * a.reps * a.streams / (ms * 1e-3) / 1e9          # GB/s in one direction

    res = {}
    for name, duplex in (("h2d_only", False), ("duplex_each_direction", True)):
        g = run(duplex)
        t = torch.tensor([g], dtype=torch.float64, device="cuda")
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            vals = [float(v.item()) for v in out]
        else:
            vals = [g]
        res[name] = {"per_gpu_gbs": [round(v, 2) for v in vals], "min": round(min(vals), 2), "mean": round(sum(vals) / len(vals), 2),
                     "max": round(max(vals), 2), "aggregate_gbs": round(sum(vals), 1)}
    if rank == 0:
        os.write(json_fd, (json.dumps({"tool": "h2d_ceiling", "n_gpus": world, "chunk_mb": a.chunk_mb, "streams_per_gpu": a.streams,
                                       "placement_rank0": placement, **res}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
