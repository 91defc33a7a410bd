#!/usr/bin/env python
"""bench.py -- MCRAW frame-decode throughput on B200 (BASELINE.json metric: decoded Mpix/s, % of HBM roofline,
next to the reference CPU decoder on the same box's host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4] [--impl ours|reference]

One step = one pass of the hot path over one batch (the whole workload clip) of synthetic frames.
  value     device-resident: compressed frames already in HBM, K batches enqueued back to back through the
            C-ABI (mcraw_decode_batch), timed with CUDA events on the launching stream, max over ranks.
  e2e       the same batch through mcraw_decode_batch_host: compressed frames in PINNED HOST memory, H2D on
            side streams overlapped with decode, per-frame results read back to the host every step.
  roofline  dominant kernel (k_units / k_legacy_decode): algorithmic bytes per launch / mean launch duration
            from CUDA events recorded around that kernel inside the timed region.
  cpu_baseline  the unmodified reference (oracle/_ref) on all host cores over a bounded sample (rank 0, N=1).
Multi-GPU: one process per GPU (torchrun), frames are independent -> every rank decodes its own clip, no
collective on the data path ("scaling": "weak"); torch.distributed is only used for the barrier / max-reduce.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "decoded_mpix_per_s"
UNIT = "Mpix/s"

WORKLOADS = {
    # name: (description, width, height, compression_type, generator, maxval, frames, distinct)
    "c2": ("c2: 240 x 1920x1080 12-bit photon clip, compressionType 7, batch decode", 1920, 1080, 7, "photon", 4095, 240, 16),
    "c1": ("c1: single 4080x3072 10-bit photon frame, compressionType 7", 4080, 3072, 7, "photon", 1023, 1, 1),
    "c3": ("c3: 4080x3072 flat+noise (0-bit / 10-bit blocks) clip, compressionType 7", 4080, 3072, 7, "flatnoise", 1023, 125, 8),
    "c4": ("c4: 4000x3000 10-bit photon legacy clip, compressionType 6", 4000, 3000, 6, "photon", 1023, 64, 8),
}


def make_streams(wl, frames_override=None):
    from motioncam_decoder_b200 import testvec as tv
    desc, w, h, ct, gen, maxval, frames, distinct = WORKLOADS[wl]
    if frames_override:
        frames = frames_override
    distinct = min(distinct, frames)
    streams = []
    for s in range(distinct):
        img = tv.gen_photon(w, h, maxval, seed=s + 1) if gen == "photon" else tv.gen_flatnoise(w, h, 256, seed=s + 1)
        streams.append(tv.encode_current(img) if ct == 7 else tv.encode_legacy(img))
    return desc, w, h, ct, frames, streams


def payload_bytes(stream, ct):
    """Bytes the dominant kernel itself has to read: payload only for type 7 (bitsOffset - 16), all for type 6."""
    if ct == 7:
        return int(np.frombuffer(stream[8:12].tobytes(), dtype="<u4")[0]) - 16
    return len(stream)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the benchmark runs."""

    def __init__(self, index, interval=0.02):
        super().__init__(daemon=True)
        self.index = index
        self.interval = interval   # NVML queries are not free for the GPU: a 4 ms poll cost ~10 % of a 0.33 ms step
        self.samples = []   # (t, sm_mhz, reasons_mask)
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(self.interval)

    def summary(self, windows):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}
        inwin = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        window = "timed regions"
        if len(inwin) < 3:
            inwin = self.samples
            window = "whole loaded run incl. warm-up (timed region too short for 3 samples)"
        mask = 0
        for s in inwin:
            mask |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in inwin), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names.items() if mask & b), "samples": len(inwin), "window": window}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    desc, w, h, ct, frames, streams = make_streams(args.workload, args.frames)
    cores = os.cpu_count() or 1
    kind, fn = cpu_bench_fn(ol)
    ins = (ctypes.c_void_p * len(streams))(*[s.ctypes.data for s in streams])
    lens = (ctypes.c_size_t * len(streams))(*[len(s) for s in streams])
    # one step = a bounded sample of the workload: whole passes over the distinct frames, as many as the workload
    # has (frames / distinct) unless that would take longer than ~0.4 s per step on this host
    done0 = ctypes.c_int64()
    if kind == "reference":
        t_pass = fn(ct, ins, lens, len(streams), w, h, cores, 1, 1, ctypes.byref(done0))
    else:
        t_pass = fn(ct, ins, lens, len(streams), w, h, cores, 1, ctypes.byref(done0))
    iters = max(1, min(frames // len(streams), int(0.4 / max(t_pass, 1e-4))))
    sample = iters * len(streams)

    def step():
        done = ctypes.c_int64()
        if kind == "reference":
            t = fn(ct, ins, lens, len(streams), w, h, cores, iters, 0, ctypes.byref(done))
        else:
            t = fn(ct, ins, lens, len(streams), w, h, cores, iters, ctypes.byref(done))
        assert done.value == sample, (done.value, sample)
        return t

    for _ in range(args.warmup):
        step()
    t = sum(step() for _ in range(args.steps))
    mpix = sample * w * h * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": desc, "frames_per_step_sample": sample, "frames_in_workload": frames, "width": w, "height": h},
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} of {frames} frames per step, {cores} threads frame-parallel, "
                                   f"g++ -O3 build of the unmodified reference" if kind == "reference" else
                                   f"{sample} of {frames} frames per step, {cores} threads, C restatement (oracle port)"},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_bench_fn(ol):
    if ol.have_ref():
        c = ctypes.CDLL(ol.REF_SO)
        f = c.mcref_bench_mt
        f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_int,
                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        f.restype = ctypes.c_double
        return "reference", f
    c = ctypes.CDLL(ol.ORACLE_SO)
    f = c.oracle_bench_mt
    f.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_int,
                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    f.restype = ctypes.c_double
    return "port", f


def cpu_baseline(streams, w, h, ct, budget_s=12.0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    kind, fn = cpu_bench_fn(ol)
    cores = os.cpu_count() or 1
    ins = (ctypes.c_void_p * len(streams))(*[s.ctypes.data for s in streams])
    lens = (ctypes.c_size_t * len(streams))(*[len(s) for s in streams])
    done = ctypes.c_int64()

    def run(iters, warm):
        if kind == "reference":
            return fn(ct, ins, lens, len(streams), w, h, cores, iters, warm, ctypes.byref(done))
        return fn(ct, ins, lens, len(streams), w, h, cores, iters, ctypes.byref(done))

    t1 = run(1, 1)
    iters = int(max(1, min(2000, budget_s / max(t1, 1e-4))))
    t = run(iters, 1)
    frames_done = done.value
    return {"value": frames_done * w * h / t / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(streams)} distinct frames x {iters} passes = {frames_done} frames in {t:.1f} s, "
                      f"{cores} threads frame-parallel, " +
                      ("unmodified reference built with g++ -O3 -include cstring" if kind == "reference" else "C restatement")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cross-batch", type=int, default=0,
                    help="experiment (profiles/README.md): CTAs the pixel kernel leaves to the next batch's index kernel; 0 = off")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    # Libraries (NCCL prints its version) write to the C-level stdout; the driver wants exactly ONE JSON line there.
    # Everything else goes to stderr: fd 1 is pointed at fd 2 and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from motioncam_decoder_b200 import capi, numa

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the decode path is CUDA only (no CPU fallback)")
    torch.cuda.set_device(local)
    props = torch.cuda.get_device_properties(local)
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        placement = numa.bind_to_gpu_node(bdf)
    except AttributeError:
        placement = "numa: torch does not expose the PCI address, not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    desc, w, h, ct, frames, streams = make_streams(args.workload, args.frames)
    ctx = capi.Context(local)
    ctx.set_kernel_timing(8)       # CUDA events around the kernels of every 8th batch inside the timed region
    stream = torch.cuda.Stream()
    sh = stream.cuda_stream

    # ---- device-resident inputs (distinct buffers per frame, contents cycle over the distinct streams)
    src_ptrs, dst_ptrs, items = [], [], []
    for i in range(frames):
        s = streams[i % len(streams)]
        sp = ctx.device_alloc(len(s) + 256)
        dp = ctx.device_alloc(w * h * 2 + 256)
        ctx.h2d(sp, s)
        src_ptrs.append(sp)
        dst_ptrs.append(dp)
        items.append((sp, len(s), w, h, ct, dp, w * h))
    descs, n = capi.Context.make_descs(items)
    comp_bytes = sum(len(streams[i % len(streams)]) for i in range(frames))
    pay_bytes = sum(payload_bytes(streams[i % len(streams)], ct) for i in range(frames))
    out_bytes = frames * w * h * 2
    pix = frames * w * h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    # ---- correctness of what is being timed: one batch checked against the expected element counts
    ctx.decode_batch(descs, n, sh)
    written, status = ctx.batch_wait(n)
    assert all(v == w * h for v in written) and not any(status), "decode failed in bench set-up"
    # set-up, not warm-up: the context keeps a few slots (scratch, descriptor tables) that are allocated on first use;
    # touch all of them now so that no cudaMalloc lands inside the timed region
    for _ in range(8):
        ctx.decode_batch(descs, n, sh)
    ctx.batch_wait(n)

    if args.cross_batch:
        # the sources of the device-resident leg are complete in HBM before any timed call: the promise this switch needs
        ctx.set_sources_resident(args.cross_batch)
    for _ in range(args.warmup):
        ctx.decode_batch(descs, n, sh)
    ctx.batch_wait(n)
    barrier()
    m0, k0, c0 = ctx.kernel_time_totals()
    l0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        ctx.decode_batch(descs, n, sh)
    ev1.record(stream)
    written, status = ctx.batch_wait(n)
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    ms = ev0.elapsed_time(ev1)
    assert all(v == w * h for v in written) and not any(status)
    m1, k1, c1 = ctx.kernel_time_totals()
    launches = ctx.kernel_launches - l0
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    per_rank_ms = [ms]
    if world > 1:
        gathered = [torch.zeros_like(tms) for _ in range(world)]
        dist.all_gather(gathered, tms)
        per_rank_ms = [float(g.item()) for g in gathered]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * pix * args.steps / (ms_max * 1e-3) / 1e6

    # ---- roofline of the dominant kernel from the events recorded around it inside the timed region
    peak, peak_src = measured_peak()
    chunks = max(1, c1 - c0)                      # TIMED launches of the dominant kernel inside the timed region
    main_ms = (k1 - k0) / chunks                  # mean duration of one launch (CUDA events around the kernel)
    meta_ms = (m1 - m0) / chunks
    alg_main = pay_bytes + out_bytes              # algorithmic bytes of one launch: the whole batch (payload + output)
    achieved = alg_main / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    step_gbs = (comp_bytes + out_bytes) * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end: pinned host inputs -> H2D on side streams -> decode -> results to host, every step
    # one pinned ring holding the clip back to back (256-byte aligned frames), as a container reader would fill it
    offs, total = [], 0
    for i in range(frames):
        offs.append(total)
        total += (len(streams[i % len(streams)]) + 255) & ~255
    ring_ptr, ring = ctx.pinned_array(total + 256)
    pinned = [ring_ptr]
    hitems = []
    for i in range(frames):
        s = streams[i % len(streams)]
        ring[offs[i]:offs[i] + len(s)] = s
        hitems.append((ring_ptr + offs[i], len(s), w, h, ct, dst_ptrs[i], w * h))
    hdescs, hn = capi.Context.make_descs(hitems)
    ctx.decode_batch_host(hdescs, hn, sh)
    written, status = ctx.batch_wait(hn)
    assert all(v == w * h for v in written) and not any(status)
    t0 = time.perf_counter()
    ctx.decode_batch_host(hdescs, hn, sh)
    ctx.batch_wait(hn)
    est = time.perf_counter() - t0
    e2e_steps = int(max(3, min(args.steps, 15.0 / max(est, 1e-4))))
    barrier()
    t_a = time.perf_counter()
    ev0.record(stream)
    for _ in range(e2e_steps):
        ctx.decode_batch_host(hdescs, hn, sh)
        written, status = ctx.batch_wait(hn)
    ev1.record(stream)
    barrier()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    e2e_ms = ev0.elapsed_time(ev1)
    assert all(v == w * h for v in written) and not any(status)
    ems = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = world * pix * e2e_steps / (float(ems.item()) * 1e-3) / 1e6

    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary(windows)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(streams, w, h, ct)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "ms_per_step_by_rank": [round(v / args.steps, 5) for v in per_rank_ms],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": {"workload": desc, "frames_per_gpu": frames, "distinct_frames": len(streams), "width": w, "height": h,
                       "compression_type": ct, "compressed_bytes_per_frame": comp_bytes / frames,
                       "compressed_bytes_per_pixel": comp_bytes / pix, "algorithmic_bytes_per_pixel": (comp_bytes + out_bytes) / pix,
                       "l2": f"every frame has its own input and output buffer: {(comp_bytes + out_bytes) / 1e6:.0f} MB touched per step "
                             f"(> 126 MB L2), no flush needed",
                       "parallelism": f"frame-parallel, {world} rank(s), no collective",
                       "cross_batch_ctas": args.cross_batch},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_units" if ct == 7 else "k_legacy_decode",
                         "algorithmic_bytes_per_launch": alg_main, "kernel_ms_per_launch": main_ms,
                         "timed_launches": chunks,
                         "index_kernels_ms_per_launch": meta_ms, "peak_source": peak_src,
                         "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak,
                                        "bytes_per_step": comp_bytes + out_bytes, "frac_of_8000": step_gbs / 8000.0}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": comp_bytes, "d2h_bytes_per_step": 16 * frames,
                    "steps": e2e_steps, "ms_per_step": float(ems.item()) / e2e_steps,
                    "h2d_gbs": comp_bytes * e2e_steps / (float(ems.item()) * 1e-3) / 1e9,
                    "path": "mcraw_decode_batch_host: pinned host ring -> staged H2D on side streams -> decode -> device u16, "
                            "per-frame results D2H", "placement": placement},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    for p in src_ptrs + dst_ptrs:
        ctx.device_free(p)
    for p in pinned:
        ctx.pinned_free(p)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
