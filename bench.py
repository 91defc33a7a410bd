#!/usr/bin/env python
"""bench.py -- MCRAW frame-decode throughput on B200 (BASELINE.json metric: decoded Mpix/s, % of HBM roofline,
next to the reference CPU decoder on the same box's host cores).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4]

With no --workload the ONE JSON line carries every BASELINE.json configuration: the headline (top-level keys) is
config 2 (c2: 240 x 1920x1080, device-resident batch decode on 1 GPU; weak-scaled at N > 1), and "workloads" holds
short legs for c1 (single 4080x3072 frame), c3 (the 1000-frame 4080x3072 flat+noise clip, split frame-parallel over
the N ranks: strong scaling), c4 (legacy 4000x3000) and c5 (.mcraw file -> pinned ring -> H2D -> decode, audio passed
through).  --workload X runs that one workload as the headline and nothing else.

One step = one pass of the hot path over one batch (the rank's whole clip) of synthetic frames.
  value         device-resident: compressed frames already in HBM, K batches enqueued back to back through the C-ABI
                (mcraw_decode_batch), CUDA events on the launching stream, max over ranks.
  e2e           the same batch through mcraw_decode_batch_host: compressed frames in PINNED HOST memory, H2D on side
                streams overlapped with decode, per-frame results read back every step; pixels stay on the device
                (the north star's "-> 16-bit buffers").  e2e_host_out: same, and the pixels come back to pinned host
                memory (the reference's own loadFrame contract), D2H overlapped with H2D + decode.
  roofline      dominant kernel: algorithmic bytes per launch / mean launch duration from CUDA events recorded around
                that kernel inside the timed region.
  pixels_verified  after EVERY timed loop all output frames are check-summed on the device (mcraw_checksum_frames)
                and compared with the checksums of the source images; outputs are poisoned before the loop, and three
                frames per leg are also copied back and compared sample by sample.
  cpu_baseline  the unmodified reference (oracle/_ref) on all host cores; same sampling rule as --impl reference
                (a step = whole passes over the distinct frames lasting >= 1 s).
Multi-GPU: one process per GPU (torchrun), frames are independent -> no collective on the data path;
torch.distributed only provides the barrier and the max-reduce of the timings.
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
KSLOTS = 6   # plan slots of a context (mcraw_capi.cu kSlots): c1 cycles over this many copies so that every call is a plan hit

WORKLOADS = {
    # name: (description, width, height, compression_type, generator, maxval, frames, distinct, strong-scaled over ranks)
    "c2": ("c2: 240 x 1920x1080 12-bit photon clip, compressionType 7, batch decode", 1920, 1080, 7, "photon", 4095, 240, 16, False),
    "c1": ("c1: single 4080x3072 10-bit photon frame, compressionType 7", 4080, 3072, 7, "photon", 1023, 1, 1, False),
    "c3": ("c3: 1000 x 4080x3072 flat+noise (0-bit / 10-bit blocks) clip, compressionType 7, frame-parallel over the ranks",
           4080, 3072, 7, "flatnoise", 1023, 1000, 8, True),
    "c4": ("c4: 64 x 4000x3000 10-bit photon legacy clip, compressionType 6", 4000, 3000, 6, "photon", 1023, 64, 8, False),
}
C5_DESC = "c5: .mcraw file (c3 frames + audio chunks) -> pread into pinned ring -> overlapped H2D -> decode -> device u16"


def make_streams(wl, want_images=True):
    from motioncam_decoder_b200 import testvec as tv
    desc, w, h, ct, gen, maxval, frames, distinct, strong = WORKLOADS[wl]
    distinct = min(distinct, frames)
    streams, images = [], []
    for s in range(distinct):
        seed = 1234 if wl == "c1" else s + 1
        img = tv.gen_photon(w, h, maxval, seed=seed) if gen == "photon" else tv.gen_flatnoise(w, h, 256, seed=seed)
        streams.append(tv.encode_current(img) if ct == 7 else tv.encode_legacy(img))
        images.append(img if want_images else None)
    return streams, images


def workload_config(wl, streams, world):
    """The workload, described identically by both arms (--impl ours / reference)."""
    desc, w, h, ct, gen, maxval, frames, distinct, strong = WORKLOADS[wl]
    comp = sum(len(streams[i % len(streams)]) for i in range(frames)) / frames
    return {"workload": desc, "frames": frames, "distinct_frames": len(streams), "width": w, "height": h, "compression_type": ct,
            "compressed_bytes_per_frame": comp, "compressed_bytes_per_pixel": comp / (w * h),
            "algorithmic_bytes_per_pixel": comp / (w * h) + 2.0,
            "l2": "every frame has its own input and output buffer; a step touches far more than the 126 MB L2 "
                  "(c1: six copies of the frame are cycled), no flush needed",
            "scaling": "strong (clip split over the ranks)" if strong else "weak (clip per rank)",
            "parallelism": f"frame-parallel, {world} rank(s), no collective"}


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
            window = "whole loaded run incl. warm-up (timed regions too short for 3 samples)"
        mask = 0
        for s in inwin:
            mask |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in inwin), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in names.items() if mask & b), "samples": len(inwin), "window": window}


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own decoder on the host cores.  ONE sampling rule for both places that use it
# (--impl reference and the cpu_baseline leg): a step = whole passes over the distinct frames lasting >= min_seconds.
# ---------------------------------------------------------------------------------------------------------------
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


class CpuReference:
    def __init__(self, streams, w, h, ct, min_seconds=1.0):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        self.kind, self.fn = cpu_bench_fn(ol)
        self.cores = os.cpu_count() or 1
        self.streams, self.w, self.h, self.ct = streams, w, h, ct
        self.ins = (ctypes.c_void_p * len(streams))(*[s.ctypes.data for s in streams])
        self.lens = (ctypes.c_size_t * len(streams))(*[len(s) for s in streams])
        t1 = self._run(1, 1)[0]                                 # one warm pass to size the step
        self.iters = int(max(1, min(100000, np.ceil(min_seconds / max(t1, 1e-5)))))

    def _run(self, iters, warm):
        done = ctypes.c_int64()
        if self.kind == "reference":
            t = self.fn(self.ct, self.ins, self.lens, len(self.streams), self.w, self.h, self.cores, iters, warm, ctypes.byref(done))
        else:
            t = self.fn(self.ct, self.ins, self.lens, len(self.streams), self.w, self.h, self.cores, iters, ctypes.byref(done))
        return t, done.value

    def step(self):
        """(seconds, frames decoded) of one step; one untimed pass inside the call warms the threads' buffers."""
        t, done = self._run(self.iters, 1)
        assert done == self.iters * len(self.streams), (done, self.iters)
        return t, done

    def measure(self, steps, warmup=1):
        for _ in range(warmup):
            self.step()
        t = frames = 0
        for _ in range(steps):
            a, b = self.step()
            t += a
            frames += b
        mpix = frames * self.w * self.h / t / 1e6
        what = "unmodified reference built with g++ -O3 -include cstring" if self.kind == "reference" else "C restatement (oracle port)"
        return {"value": mpix, "unit": UNIT, "cores": self.cores, "kind": self.kind, "seconds_timed": t,
                "sample": f"{steps} step(s) x {self.iters} passes over {len(self.streams)} distinct frames = {frames} frames in "
                          f"{t:.2f} s, {self.cores} threads frame-parallel, {what}"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    wl = args.workload or "c2"
    streams, _ = make_streams(wl, want_images=False)
    desc, w, h, ct = WORKLOADS[wl][:4]
    ref = CpuReference(streams, w, h, ct)
    res = ref.measure(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds_timed"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": workload_config(wl, streams, world),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.workload:
        # the other configurations, one step each (same rule)
        others = {}
        for o in ("c1", "c3", "c4"):
            st, _ = make_streams(o, want_images=False)
            r = CpuReference(st, *WORKLOADS[o][1:4]).measure(1, 1)
            others[o] = {"value": r["value"], "unit": UNIT, "sample": r["sample"]}
        line["workloads"] = others
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """Per-process plumbing shared by the legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from motioncam_decoder_b200 import capi, numa
        self.torch, self.dist, self.capi = torch, dist, capi
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the decode path is CUDA only (no CPU fallback)")
        torch.cuda.set_device(self.local)
        props = torch.cuda.get_device_properties(self.local)
        try:
            bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
            self.placement = numa.bind_to_gpu_node(bdf)
        except AttributeError:
            self.placement = "numa: torch does not expose the PCI address, not bound"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.ctx = capi.Context(self.local)
        self.stream = torch.cuda.Stream()
        self.sh = self.stream.cuda_stream
        self.windows = []
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        if self.world == 1:
            return [float(v)]
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(g.item()) for g in out]


def h2d_ceiling(env, chunk_bytes=96 << 20, reps=8):
    """Pinned-host -> device copy bandwidth of THIS rank while every rank does the same: plain cudaMemcpyAsync of the
    chunk size mcraw_decode_batch_host uses, nothing else on the GPU.  The ceiling of the e2e legs."""
    torch = env.torch
    host = torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(chunk_bytes, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(env.stream):
        dev.copy_(host, non_blocking=True)
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(env.stream):
        e0.record(env.stream)
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        e1.record(env.stream)
    env.barrier()
    ms = e0.elapsed_time(e1)
    return chunk_bytes * reps / (ms * 1e-3) / 1e9


def duplex_ceiling(env, in_bytes=96 << 20, out_bytes=200 << 20, reps=6):
    """Both directions of the link at once, nothing else on the GPU: pinned host -> device in 96 MB chunks on one stream while
    device -> pinned host runs on another (about two output bytes per input byte, the C2 mix), all ranks together.  The
    ceiling of the host-out legs: what the platform gives each direction while the other one is busy."""
    torch = env.torch
    hin = torch.empty(in_bytes, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")
    hout = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    dout = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for timed in (False, True):
        torch.cuda.synchronize()
        env.barrier()
        with torch.cuda.stream(s_in):
            ev[0].record(s_in)
            for _ in range(reps if timed else 1):
                din.copy_(hin, non_blocking=True)
            ev[1].record(s_in)
        with torch.cuda.stream(s_out):
            ev[2].record(s_out)
            for _ in range(reps if timed else 1):
                hout.copy_(dout, non_blocking=True)
            ev[3].record(s_out)
        torch.cuda.synchronize()
    env.barrier()
    return {"h2d_gbs": in_bytes * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9,
            "d2h_gbs": out_bytes * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9,
            "how": f"{reps} x {in_bytes >> 20} MB H2D and {reps} x {out_bytes >> 20} MB D2H, pinned, on two streams at the same time"}


def run_leg(env, wl, steps, warmup, headline):
    """All measurements of one workload on this rank.  Returns the dict that goes into the JSON line."""
    torch, capi, ctx = env.torch, env.capi, env.ctx
    from motioncam_decoder_b200 import shard
    desc, w, h, ct, gen, maxval, frames_total, distinct, strong = WORKLOADS[wl]
    streams, images = make_streams(wl)
    expect = [capi.checksum_u16(im) for im in images]
    if strong:
        mine = list(shard.shard_contiguous(frames_total, env.world, env.rank))      # global frame indices of this rank
    else:
        mine = list(range(frames_total))
    copies = KSLOTS if frames_total == 1 else 1       # c1: cycle over that many copies of the frame (> L2, every call a plan hit)
    nf = len(mine) * copies
    gidx = [mine[i % len(mine)] for i in range(nf)] if mine else []
    npix = w * h
    in_pitch = (max(len(s) for s in streams) + 511) & ~255
    out_pitch = (npix * 2 + 511) & ~255
    src = torch.empty(max(nf, 1) * in_pitch, dtype=torch.uint8, device="cuda")
    dst = torch.empty(max(nf, 1) * out_pitch, dtype=torch.uint8, device="cuda")
    dev_streams = [torch.from_numpy(s).cuda() for s in streams]
    items = []
    for i in range(nf):
        k = gidx[i] % len(streams)
        src[i * in_pitch:i * in_pitch + len(streams[k])].copy_(dev_streams[k])
        items.append((src.data_ptr() + i * in_pitch, len(streams[k]), w, h, ct, dst.data_ptr() + i * out_pitch, npix))
    torch.cuda.synchronize()
    per_step = len(mine)                               # frames of one step (one batch) on this rank
    batches = [capi.Context.make_descs(items[c * per_step:(c + 1) * per_step]) for c in range(copies)] if mine else []
    dst_ptrs = [it[5] for it in items]
    comp_step = sum(len(streams[g % len(streams)]) for g in mine)
    pay_step = sum(payload_bytes(streams[g % len(streams)], ct) for g in mine)
    out_step = len(mine) * npix * 2
    verified = {"checks": 0, "frames": 0, "ok": True}

    def poison():
        dst.fill_(0xA5)
        torch.cuda.synchronize()

    def verify(ptrs, idx, what):
        """Every output frame against the checksum of its source image (computed on the device), three of them also sample
        by sample.  Raises on a mismatch: a wrong pixel must never turn into a throughput number."""
        if not ptrs:
            return
        sums = ctx.checksum_frames(ptrs, [npix] * len(ptrs))
        bad = [i for i, s in enumerate(sums) if s != expect[idx[i] % len(streams)]]
        out = np.empty((h, w), dtype=np.uint16)
        for i in sorted({0, len(ptrs) // 2, len(ptrs) - 1}):
            ctx.d2h(out, ptrs[i])
            if not np.array_equal(out, images[idx[i] % len(streams)]):
                bad.append(i)
        verified["checks"] += 1
        verified["frames"] += len(ptrs)
        if bad:
            verified["ok"] = False
            raise SystemExit(f"bench.py: {wl} {what}: {len(bad)} of {len(ptrs)} decoded frames differ from their source images "
                             f"(first: {bad[:8]})")

    def run_batches(count, fn):
        for i in range(count):
            d, n = batches[i % copies]
            fn(d, n)

    res = {"frames_per_step_this_rank": per_step}
    # kernel-timing events on a SAMPLE of the timed region's batches (four of a hundred, two of twenty): a sampled batch runs its
    # two kernels one after the other between events, and it and the batch behind it leave the chain of programmatic launches
    ctx.set_kernel_timing(steps // 4 if steps > 40 else max(2, steps // 2))     # any window of `steps` batches holds at least one sample
    if mine:
        # set-up, not warm-up: the context's plan slots are allocated on first use; touch all of them now
        poison()
        run_batches(2 * KSLOTS, lambda d, n: ctx.decode_batch(d, n, env.sh))
        written, status = ctx.batch_wait(per_step)
        assert all(v == npix for v in written) and not any(status), f"{wl}: decode failed in bench set-up"
        verify(dst_ptrs, gidx, "set-up")
        poison()
        run_batches(max(warmup, copies), lambda d, n: ctx.decode_batch(d, n, env.sh))
        ctx.batch_wait(per_step)
        poison()
    env.barrier()
    m0, k0, c0 = ctx.kernel_time_totals()
    l0 = ctx.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_a = time.perf_counter()
    ev0.record(env.stream)
    if mine:
        run_batches(steps, lambda d, n: ctx.decode_batch(d, n, env.sh))
    ev1.record(env.stream)
    if mine:
        written, status = ctx.batch_wait(per_step)
    env.barrier()
    t_b = time.perf_counter()
    env.windows.append((t_a, t_b))
    ms = ev0.elapsed_time(ev1)
    if mine:
        assert all(v == npix for v in written) and not any(status)
        ncheck = min(nf, steps * per_step)                                       # frames the timed loop has written
        verify(dst_ptrs[:ncheck], gidx[:ncheck], "device-resident timed loop")
    m1, k1, c1 = ctx.kernel_time_totals()
    res["gpu_launches"] = ctx.kernel_launches - l0
    per_rank_ms = env.gather(ms)
    ms_max = max(per_rank_ms)
    total_frames = env.sum_over_ranks(per_step)
    res["ms_per_step"] = ms_max / steps
    res["ms_per_step_by_rank"] = [round(v / steps, 5) for v in per_rank_ms]
    res["value"] = total_frames * npix * steps / (ms_max * 1e-3) / 1e6
    res["steps"] = steps

    # ---- roofline of the dominant kernel from the events recorded around it inside the timed region
    peak, peak_src = measured_peak()
    chunks = max(1, c1 - c0)
    main_ms = (k1 - k0) / chunks
    meta_ms = (m1 - m0) / chunks
    alg_main = pay_step + out_step
    achieved = alg_main / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(wl, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    if per_step != frames_total:          # a strong-scaled leg: the capture was one launch over the WHOLE clip
        traffic = None
    step_gbs = (comp_step + out_step) * steps / (ms * 1e-3) / 1e9 if mine else 0.0
    res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                       "traffic": traffic, "kernel": "k_units" if ct == 7 else "k_legacy_warp",
                       "algorithmic_bytes_per_launch": alg_main, "kernel_ms_per_launch": main_ms, "timed_launches": chunks,
                       "index_kernels_ms_per_launch": meta_ms, "peak_source": peak_src,
                       "whole_step": {"achieved": step_gbs, "frac": step_gbs / peak, "bytes_per_step": comp_step + out_step,
                                      "frac_of_8000": step_gbs / 8000.0, "note": "this rank's bytes / this rank's time"}}
    ctx.set_kernel_timing(0)

    # ---- cold plans: every call presents descriptors the context has not seen (validation, work list, plan upload inside).
    #      Two flavours: consecutive batches write ALTERNATING output areas (a double-buffered consumer: the new batch may be
    #      chained to the one before, its plan goes up beside the decode stream), and all batches write the SAME buffers (no
    #      chain: a batch with other descriptors may not overlap one that still writes there).
    if headline and mine and copies == 1:
        def cold_loop(plans):
            for d, n in plans:
                ctx.decode_batch(d, n, env.sh)
            ctx.batch_wait(per_step)
            env.barrier()
            ev0.record(env.stream)
            csteps = 3 * len(plans)
            for i in range(csteps):
                d, n = plans[i % len(plans)]
                ctx.decode_batch(d, n, env.sh)
            ev1.record(env.stream)
            written, status = ctx.batch_wait(per_step)
            assert all(v == npix for v in written) and not any(status)
            env.barrier()
            return env.max_over_ranks(ev0.elapsed_time(ev1)) / csteps

        dst2 = torch.empty_like(dst)
        shift = dst2.data_ptr() - dst.data_ptr()
        alt = [capi.Context.make_descs([it[:5] + (it[5] + (shift if k % 2 else 0), it[6] + 8 * (k + 1)) for it in items]) for k in range(KSLOTS + 2)]
        poison()
        dst2.fill_(0xA5)
        res["cold_plan_ms_per_step"] = cold_loop(alt)
        verify(dst_ptrs, gidx, "cold plans, first output area")
        verify([p_ + shift for p_ in dst_ptrs], gidx, "cold plans, second output area")
        del dst2
        same = [capi.Context.make_descs([it[:6] + (it[6] + 8 * (k + 1),) for it in items]) for k in range(KSLOTS + 1)]
        res["cold_plan_same_outputs_ms_per_step"] = cold_loop(same)

    # ---- end to end: pinned host inputs -> H2D on side streams -> decode -> results to host, every step.
    #      One pinned ring holds the clip back to back (256-byte aligned frames), as a container reader would fill it;
    #      big clips use a bounded prefix of the rank's shard (stated).
    e2e_n = per_step
    budget = 1200 << 20
    while e2e_n > 1 and sum(len(streams[g % len(streams)]) for g in mine[:e2e_n]) > budget:
        e2e_n //= 2
    res["e2e"] = res["e2e_host_out"] = None
    if mine:
        offs, total = [], 0
        for g in mine[:e2e_n]:
            offs.append(total)
            total += (len(streams[g % len(streams)]) + 255) & ~255
        ring_ptr, ring = ctx.pinned_array(total + 256)
        hitems = []
        for i, g in enumerate(mine[:e2e_n]):
            s = streams[g % len(streams)]
            ring[offs[i]:offs[i] + len(s)] = s
            hitems.append((ring_ptr + offs[i], len(s), w, h, ct, dst_ptrs[i], npix))
        hdescs, hn = capi.Context.make_descs(hitems)
        e2e_comp = sum(it[1] for it in hitems)
        out_ring_ptr, _ = ctx.pinned_array(e2e_n * out_pitch + 256)
        host_out = (ctypes.c_void_p * e2e_n)(*[out_ring_ptr + i * out_pitch for i in range(e2e_n)])

        def e2e_leg(host_pixels):
            def once():
                if host_pixels:
                    ctx.decode_batch_host_out(hdescs, host_out, hn, env.sh)
                else:
                    ctx.decode_batch_host(hdescs, hn, env.sh)
                return ctx.batch_wait(hn)
            poison()
            written, status = once()
            assert all(v == npix for v in written) and not any(status)
            t0 = time.perf_counter()
            once()
            est = time.perf_counter() - t0
            n_steps = int(max(3, min(steps, 6.0 / max(est, 1e-4))))
            poison()
            env.barrier()
            t_a = time.perf_counter()
            ev0.record(env.stream)
            for _ in range(n_steps):
                written, status = once()
            ev1.record(env.stream)
            env.barrier()
            t_b = time.perf_counter()
            env.windows.append((t_a, t_b))
            wall_ms = (t_b - t_a) * 1e3
            ems = env.max_over_ranks(max(ev0.elapsed_time(ev1), 0.0))
            wall_ms = env.max_over_ranks(wall_ms)
            assert all(v == npix for v in written) and not any(status)
            verify(dst_ptrs[:e2e_n], gidx[:e2e_n], "e2e host-out leg" if host_pixels else "e2e leg")
            if host_pixels:                 # and the pixels that came back to the host
                for i in sorted({0, e2e_n // 2, e2e_n - 1}):
                    got = np.frombuffer((ctypes.c_uint8 * (npix * 2)).from_address(out_ring_ptr + i * out_pitch), dtype=np.uint16)
                    if not np.array_equal(got.reshape(h, w), images[gidx[i] % len(streams)]):
                        raise SystemExit(f"bench.py: {wl} e2e host-out: host frame {i} differs from its source image")
            # the host-out leg ends when the last D2H has landed, which the stream events do not see: wall clock there
            t_ms = wall_ms if host_pixels else ems
            tot = env.sum_over_ranks(e2e_n)
            return {"value": tot * npix * n_steps / (t_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": e2e_comp,
                    "d2h_bytes_per_step": 16 * e2e_n + (e2e_n * npix * 2 if host_pixels else 0), "steps": n_steps,
                    "frames_per_step_this_rank": e2e_n, "ms_per_step": t_ms / n_steps,
                    "h2d_gbs": e2e_comp * n_steps / (t_ms * 1e-3) / 1e9,
                    "timed_with": "wall clock between barriers (max over ranks)" if host_pixels else "CUDA events on the decode stream (max over ranks)"}

        res["e2e"] = e2e_leg(False)
        res["e2e"]["path"] = ("mcraw_decode_batch_host: pinned host ring -> staged H2D on side streams -> decode -> device u16 "
                              "(pixels stay on the device: the consumer is another GPU stage), per-frame results D2H")
        res["e2e"]["placement"] = env.placement
        res["e2e_host_out"] = e2e_leg(True)
        res["e2e_host_out"]["path"] = ("mcraw_decode_batch_host_out: as e2e, and every decoded frame is copied to pinned host memory "
                                       "(the reference's loadFrame contract), D2H of chunk c overlapping H2D + decode of chunk c+1")
        ctx.pinned_free(ring_ptr)
        ctx.pinned_free(out_ring_ptr)
    res["pixels_verified"] = bool(verified["ok"] and (verified["checks"] >= 4 or not mine))
    res["verified_frames_total"] = verified["frames"]
    del src, dst, dev_streams
    torch.cuda.empty_cache()
    return res, streams


def run_file_leg(env, frames=64, reps=3):
    """BASELINE config 5: a .mcraw file in tmpfs -> motioncam::Decoder::loadFramesToDevice (pread into the pinned ring, staged
    H2D, decode) on every rank's shard of the timestamp-sorted frame list; audio chunks passed through on the host."""
    torch, capi, ctx = env.torch, env.capi, env.ctx
    from motioncam_decoder_b200 import hostapi, shard, testvec as tv
    os.environ["MCRAW_B200_DEVICE"] = str(env.local)
    w, h = 4080, 3072
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(d, f"mcraw_bench_c5_{os.getpid() if env.world == 1 else 'shared'}.mcraw")
    rng = np.random.default_rng(7)
    audio = [(1_000_000 * i if i % 2 == 0 else None, rng.integers(-3000, 3000, 1920 * 2, dtype=np.int16)) for i in range(16)]
    distinct = 4
    images = [tv.gen_flatnoise(w, h, 256, seed=s + 1) for s in range(distinct)]
    expect = [capi.checksum_u16(im) for im in images]
    if env.rank == 0:
        streams = [tv.encode_current(im) for im in images]
        fr = [{"timestamp": 1000 + 33 * i, "data": streams[i % distinct], "width": w, "height": h, "compressionType": 7}
              for i in range(frames)]
        tv.write_mcraw(path, fr, audio)
    env.barrier()
    dec = hostapi.Decoder(path)
    stamps = dec.get_frames()
    idx = list(shard.shard_contiguous(len(stamps), env.world, env.rank))
    mine = [stamps[i] for i in idx]
    out_pitch = (w * h * 2 + 511) & ~255
    dst = torch.empty(max(1, len(mine)) * out_pitch, dtype=torch.uint8, device="cuda")
    ptrs = [dst.data_ptr() + i * out_pitch for i in range(len(mine))]
    caps = [w * h] * len(mine)
    if mine:
        dec.load_frames_to_device(mine, ptrs, caps)         # warm-up: ring allocation, page cache
        dst.fill_(0xA5)
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mine:
            dec.load_frames_to_device(mine, ptrs, caps)
    torch.cuda.synchronize()
    dt = env.max_over_ranks(time.perf_counter() - t0)
    env.windows.append((t0, time.perf_counter()))
    ok = True
    if mine:
        sums = ctx.checksum_frames(ptrs, caps)
        ok = all(s == expect[i % distinct] for s, i in zip(sums, idx))
    audio_ok = None
    if env.rank == 0:
        got = dec.load_audio()
        audio_ok = [(t, a.tobytes()) for t, a in got] == [(-1 if t is None else t, np.asarray(a).tobytes()) for t, a in audio]
    feed = dec.feed_description() if hasattr(dec, "feed_description") else None
    size = os.path.getsize(path)
    dec.close()
    env.barrier()
    if env.rank == 0:
        try:
            os.remove(path)
        except OSError:
            pass
    if not ok:
        raise SystemExit("bench.py: c5: frames decoded from the file differ from their source images")
    return {"workload": C5_DESC, "value": frames * w * h * reps / dt / 1e6, "unit": UNIT, "frames": frames, "reps": reps,
            "ms_per_step": 1e3 * dt / reps, "file_bytes": size, "file_gb_per_s": size * reps / dt / 1e9, "pixels_verified": bool(ok),
            "audio_ok": audio_ok, "feed": feed, "scaling": "strong (file's frame list split over the ranks)",
            "timed_with": "wall clock around Decoder::loadFramesToDevice + device sync (max over ranks)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="run only this workload as the headline (default: c2 headline + c1/c3/c4/c5 legs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    # Libraries (NCCL prints its version) write to the C-level stdout; the driver wants exactly ONE JSON line there.
    # Everything else goes to stderr: fd 1 is pointed at fd 2 and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    env = Env(args)
    sampler = ClockSampler(env.local)
    sampler.start()
    head_wl = args.workload or "c2"
    head, head_streams = run_leg(env, head_wl, args.steps, args.warmup, headline=True)
    peak_h2d = h2d_ceiling(env)
    duplex = duplex_ceiling(env)

    def ceilings(leg):
        for key in ("e2e", "e2e_host_out"):
            if leg.get(key):
                leg[key]["h2d_peak_gbs"] = peak_h2d
                leg[key]["frac_of_h2d"] = leg[key]["h2d_gbs"] / peak_h2d
        ho = leg.get("e2e_host_out")
        if ho:                                                   # host-out: the device -> host direction is the busy one
            ho["d2h_gbs"] = ho["d2h_bytes_per_step"] / (ho["ms_per_step"] * 1e-3) / 1e9
            ho["duplex_peak"] = duplex
            ho["frac_of_duplex_d2h"] = ho["d2h_gbs"] / duplex["d2h_gbs"]

    ceilings(head)
    extras = {}
    if not args.workload:
        for wl in ("c1", "c3", "c4"):
            leg, st = run_leg(env, wl, max(3, min(args.steps, 20)), args.warmup, headline=False)
            leg["config"] = workload_config(wl, st, env.world)
            ceilings(leg)
            extras[wl] = leg
        extras["c5"] = run_file_leg(env)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary(env.windows)

    cpu = None
    if env.rank == 0 and env.world == 1 and not args.no_cpu_baseline:
        from motioncam_decoder_b200 import numa
        numa.unbind()                                           # the CPU baseline gets every host core back
        desc, w, h, ct = WORKLOADS[head_wl][:4]
        r = CpuReference(head_streams, w, h, ct).measure(3, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if env.rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": env.world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "ms_per_step_by_rank": head["ms_per_step_by_rank"],
            "higher_is_better": True, "scaling": "strong" if WORKLOADS[head_wl][8] else "weak", "vs_baseline": None,
            "dtype": "u16", "data": "synthetic",
            "config": workload_config(head_wl, head_streams, env.world),
            "roofline": head["roofline"],
            "e2e": head["e2e"], "e2e_host_out": head["e2e_host_out"],
            "pixels_verified": head["pixels_verified"], "verified_frames_total": head["verified_frames_total"],
            "cold_plan_ms_per_step": head.get("cold_plan_ms_per_step"),
            "cold_plan_same_outputs_ms_per_step": head.get("cold_plan_same_outputs_ms_per_step"),
            "chain": os.environ.get("MCRAW_CHAIN", "default (24 CTAs held back for the next batch's index kernel)"),
            "gpu_launches": head["gpu_launches"],
            "clocks": clocks,
        }
        if extras:
            line["workloads"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    env.ctx.close()
    if env.world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
