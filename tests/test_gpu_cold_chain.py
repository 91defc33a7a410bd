"""GPU: batches whose descriptors are NEW every time, enqueued back to back without waiting.  Such a batch has its plan
(descriptors + work list) uploaded on a side stream while its kernels -- chained to the batch before by programmatic
launches when the outputs do not overlap -- wait for the plan's ready word themselves, and the per-frame done words carry
the slot's launch epoch instead of being zeroed.  Every frame of every batch against the source image (= the oracle's
output, checked once), for the warp index kernel (batches), the CTA index kernels (a few frames) and with the side-stream
upload switched off (MCRAW_PLAN_SIDE=0).  Reference: RawData.cpp:528-612 per frame; Decoder.cpp:184-235 is the caller."""
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _ctx_with(**env):
    from motioncam_decoder_b200 import capi
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return capi.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def _pool():
    """A few encoded frames of two sizes (checked against the oracle once)."""
    from motioncam_decoder_b200 import testvec as tv
    pool = []
    for k, (w, h) in enumerate([(640, 64), (1000, 32), (640, 64), (328, 48), (1928, 16), (640, 64)]):
        img = tv.gen_photon(w, h, 4095 if k % 2 else 1023, seed=300 + k)
        s = tv.encode_current(img, policy=tv.POLICY_ALIASES if k % 3 else tv.POLICY_MINIMAL, seed=k)
        if k == 3:
            s = tv.pad_meta_current(s, 3, 1)
        n, want = ol.oracle_decode(s, w, h)
        assert n == w * h and np.array_equal(want, img)
        pool.append((s, w, h, img))
    return pool


class _Batch:
    """n frames picked from the pool, sources and outputs in ONE device allocation each (outputs of different batches are
    disjoint address ranges: what the chain needs for batches that are not the same)."""

    def __init__(self, ctx, pool, picks):
        from motioncam_decoder_b200 import capi
        self.ctx, self.frames = ctx, [pool[p] for p in picks]
        src_off, dst_off, so, do = [], [], 0, 0
        for s, w, h, _ in self.frames:
            src_off.append(so); so += (len(s) + 255) & ~255
            dst_off.append(do); do += (2 * w * h + 255) & ~255
        self.src = ctx.device_alloc(so + 256)
        self.dst = ctx.device_alloc(do + 256)
        self.dst_bytes = do
        items = []
        for (s, w, h, _), a, b in zip(self.frames, src_off, dst_off):
            ctx.h2d(self.src + a, s)
            items.append((self.src + a, len(s), w, h, capi.COMPRESSION_CURRENT, self.dst + b, w * h, capi.frame_encoded_width(s, w, h)))
        self.dst_off = dst_off
        self.descs, self.n = capi.Context.make_descs(items)

    def poison(self):
        self.ctx.h2d(self.dst, np.full(self.dst_bytes // 2, 0xA5A5, dtype=np.uint16))

    def check(self, label):
        for i, ((s, w, h, img), b) in enumerate(zip(self.frames, self.dst_off)):
            out = np.empty(w * h, dtype=np.uint16)
            self.ctx.d2h(out, self.dst + b)
            assert np.array_equal(out.reshape(h, w), img), f"{label}: frame {i} ({w}x{h}) differs"

    def free(self):
        self.ctx.device_free(self.src)
        self.ctx.device_free(self.dst)


def _run_rounds(ctx, sizes, rounds, label):
    pool = _pool()
    rng = np.random.default_rng(7)
    batches = [_Batch(ctx, pool, rng.integers(0, len(pool), n)) for n in sizes]
    for r in range(rounds):
        for b in batches:
            b.poison()
        order = rng.permutation(len(batches)) if r else np.arange(len(batches))
        for k in order:                                    # back to back: nothing waits in between
            ctx.decode_batch(batches[k].descs, batches[k].n)
        last = batches[order[-1]]
        written, status = ctx.batch_wait(last.n)
        assert not any(status) and all(wr == f[1] * f[2] for wr, f in zip(written, last.frames)), (label, r)
        for k, b in enumerate(batches):
            b.check(f"{label} round {r} batch {k}")
    for b in batches:
        b.free()


def test_new_batches_back_to_back_warp_kernel():
    """Eleven distinct batches of 80 ... 130 frames through six plan slots: every enqueue is a plan miss, chained, k_meta_warp."""
    from motioncam_decoder_b200 import capi
    ctx = capi.Context(0)
    _run_rounds(ctx, [80, 96, 130, 81, 100, 90, 85, 120, 99, 83, 111], 3, "default")
    ctx.close()


def test_new_batches_back_to_back_few_frames():
    """Small batches: the CTA index kernels behind the same side-stream upload; mixed with big ones and repeated ones."""
    from motioncam_decoder_b200 import capi
    ctx = capi.Context(0)
    _run_rounds(ctx, [3, 80, 1, 17, 96, 40, 2, 85, 9], 3, "few")
    _run_rounds(ctx, [12, 12, 90, 90], 4, "repeats")         # fewer batches than slots: hits after the first round
    ctx.close()


def test_new_batches_without_side_upload_and_forced_warp_kernel():
    for env, label in (({"MCRAW_PLAN_SIDE": 0}, "in-stream upload"), ({"MCRAW_META_WARP": 2}, "warp kernel forced"),
                       ({"MCRAW_CHAIN": 0}, "no chain")):
        ctx = _ctx_with(**env)
        _run_rounds(ctx, [80, 5, 96, 1, 88, 30, 101, 84], 2, label)
        ctx.close()


def test_failed_frame_in_a_chained_new_batch():
    """A truncated frame inside one of the back-to-back batches: that frame fails, every other frame of every batch decodes."""
    from motioncam_decoder_b200 import capi
    ctx = capi.Context(0)
    pool = _pool()
    s0, w0, h0, img0 = pool[0]
    bad = (s0[: len(s0) // 2].copy(), w0, h0, None)
    pool_bad = pool + [bad]
    batches = [_Batch(ctx, pool_bad, [0, 1, 2] * 30), _Batch(ctx, pool_bad, [1] * 40 + [len(pool)] + [2] * 45),
               _Batch(ctx, pool_bad, [4, 5] * 44)]
    for b in batches:
        b.poison()
    for b in batches:
        ctx.decode_batch(b.descs, b.n)
    ctx.batch_wait(batches[-1].n)
    for k, b in enumerate(batches):
        for i, ((s, w, h, img), off) in enumerate(zip(b.frames, b.dst_off)):
            out = np.empty(w * h, dtype=np.uint16)
            ctx.d2h(out, b.dst + off)
            if img is None:
                assert np.all(out == 0xA5A5), "a failed frame writes nothing"
            else:
                assert np.array_equal(out.reshape(h, w), img), (k, i)
    ctx.decode_batch(batches[1].descs, batches[1].n)         # and its verdict, asked for on its own
    written, status = ctx.batch_wait(batches[1].n)
    assert written[40] == 0 and status[40] != 0 and all(wr > 0 for i, wr in enumerate(written) if i != 40)
    for b in batches:
        b.free()
    ctx.close()
