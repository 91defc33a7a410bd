"""ctypes binding of the C-ABI in include/mcraw_b200.h (libmcraw_b200.so).

This is what tests/ and bench.py call; it adds nothing to the decode path.  Loading fails loudly when the
library is not built, and Context() raises when no CUDA device is usable -- there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib

COMPRESSION_LEGACY = 6
COMPRESSION_CURRENT = 7

FRAME_OK = 0
FRAME_BAD_HEADER = 1
FRAME_TRUNCATED = 2
FRAME_BAD_BITS = 4
FRAME_BAD_META_COUNT = 8
FRAME_GEOMETRY = 16
FRAME_BAD_TYPE = 32


class FrameDesc(ctypes.Structure):
    _fields_ = [
        ("src", ctypes.c_void_p),
        ("len", ctypes.c_uint64),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("compression_type", ctypes.c_int32),
        ("encoded_width", ctypes.c_int32),
        ("dst", ctypes.c_void_p),
        ("dst_capacity_elems", ctypes.c_uint64),
    ]


OUT_RAW, OUT_BLACK_SUB, OUT_NORM_F16 = 0, 1, 2


class Levels(ctypes.Structure):
    _fields_ = [("black", ctypes.c_float * 4), ("white", ctypes.c_float), ("mode", ctypes.c_uint32)]


def apply_levels(img, black, white, mode):
    """numpy restatement of the fused epilogue (include/mcraw_b200.h, MCRAW_OUT_*) on a decoded (height, width) uint16 image:
    returns uint16 for OUT_BLACK_SUB and the BIT PATTERNS (uint16 view) of the IEEE halves for OUT_NORM_F16."""
    img = np.ascontiguousarray(img, dtype=np.uint16)
    h, w = img.shape
    cfa = (np.arange(h)[:, None] & 1) * 2 + (np.arange(w)[None, :] & 1)
    if mode == OUT_RAW:
        return img
    if mode == OUT_BLACK_SUB:
        rint = lambda v: int(min(65535, max(0, np.rint(np.float32(v)))))     # noqa: E731  (lrintf: round half to even)
        b = np.array([rint(v) for v in black], dtype=np.int64)[cfa]
        rng = np.maximum(rint(white) - b, 0)
        return np.minimum(np.maximum(img.astype(np.int64) - b, 0), rng).astype(np.uint16)
    bl = np.array(black, dtype=np.float32)
    wh = np.float32(white)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(wh > bl, np.float32(1.0) / (wh - bl), np.float32(0.0)).astype(np.float32)
    x = (img.astype(np.float32) - bl[cfa]) * scale[cfa]          # float32 throughout, like the kernel
    x = np.minimum(np.maximum(x, np.float32(0.0)), np.float32(1.0)) + np.float32(0.0)      # saturation yields +0.0, never -0.0
    return x.astype(np.float16).view(np.uint16)


EXPORTED = [
    "mcraw_version", "mcraw_ctx_create", "mcraw_ctx_destroy", "mcraw_last_error", "mcraw_ctx_device",
    "mcraw_decode_batch", "mcraw_decode_batch_host", "mcraw_batch_wait", "mcraw_decode_host",
    "mcraw_device_alloc", "mcraw_device_free", "mcraw_host_alloc_pinned", "mcraw_host_free_pinned",
    "mcraw_memcpy_h2d", "mcraw_memcpy_d2h", "mcraw_stream_sync", "mcraw_kernel_launches",
    "mcraw_last_batch_kernel_ms", "mcraw_kernel_time_totals", "mcraw_set_kernel_timing",
    "mcraw_host_register", "mcraw_host_unregister", "mcraw_frame_encoded_width",
    "mcraw_checksum_frames", "mcraw_decode_batch_host_out",
    "mcraw_batch_begin", "mcraw_batch_append_host", "mcraw_decode_batch_levels",
]

_c = None


def lib():
    global _c
    if _c is None:
        c = _lib.load(_lib.LIB_CAPI)
        vp, u64, u32, sz = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_size_t
        c.mcraw_version.restype = ctypes.c_char_p
        c.mcraw_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
        c.mcraw_ctx_destroy.argtypes = [vp]
        c.mcraw_ctx_destroy.restype = None
        c.mcraw_last_error.argtypes = [vp]
        c.mcraw_last_error.restype = ctypes.c_char_p
        c.mcraw_ctx_device.argtypes = [vp]
        c.mcraw_decode_batch.argtypes = [vp, ctypes.POINTER(FrameDesc), u32, vp]
        c.mcraw_decode_batch_host.argtypes = [vp, ctypes.POINTER(FrameDesc), u32, vp]
        c.mcraw_decode_batch_host_out.argtypes = [vp, ctypes.POINTER(FrameDesc), ctypes.POINTER(vp), u32, vp]
        c.mcraw_decode_batch_levels.argtypes = [vp, ctypes.POINTER(FrameDesc), ctypes.POINTER(Levels), u32, vp]
        c.mcraw_batch_wait.argtypes = [vp, ctypes.POINTER(u64), ctypes.POINTER(u32), u32]
        c.mcraw_decode_host.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, vp, sz, ctypes.c_int]
        c.mcraw_decode_host.restype = sz
        c.mcraw_device_alloc.argtypes = [vp, sz, ctypes.POINTER(vp)]
        c.mcraw_device_free.argtypes = [vp, vp]
        c.mcraw_host_alloc_pinned.argtypes = [vp, sz, ctypes.POINTER(vp)]
        c.mcraw_host_free_pinned.argtypes = [vp, vp]
        c.mcraw_host_register.argtypes = [vp, vp, sz, ctypes.c_int]
        c.mcraw_host_unregister.argtypes = [vp, vp]
        c.mcraw_memcpy_h2d.argtypes = [vp, vp, vp, sz, vp]
        c.mcraw_memcpy_d2h.argtypes = [vp, vp, vp, sz, vp]
        c.mcraw_stream_sync.argtypes = [vp, vp]
        c.mcraw_kernel_launches.argtypes = [vp]
        c.mcraw_kernel_launches.restype = u64
        c.mcraw_last_batch_kernel_ms.argtypes = [vp]
        c.mcraw_last_batch_kernel_ms.restype = ctypes.c_float
        c.mcraw_kernel_time_totals.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                               ctypes.POINTER(u64)]
        c.mcraw_set_kernel_timing.argtypes = [vp, u32]
        c.mcraw_frame_encoded_width.argtypes = [vp, u64, ctypes.c_int32, ctypes.c_int32]
        c.mcraw_frame_encoded_width.restype = ctypes.c_int32
        c.mcraw_checksum_frames.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u64), u32, ctypes.POINTER(u64), vp]
        _c = c
    return _c


def frame_encoded_width(stream, width, height):
    """mcraw_frame_encoded_width: what a caller who uploads a compressionType 7 frame puts into FrameDesc.encoded_width."""
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    return int(lib().mcraw_frame_encoded_width(stream.ctypes.data, stream.size, width, height))


CHECKSUM_K = 0x9E3779B97F4A7C15


def checksum_u16(img):
    """The checksum of mcraw_checksum_frames, restated with numpy (uint64 arithmetic wraps mod 2^64)."""
    v = np.ascontiguousarray(img, dtype=np.uint16).reshape(-1).astype(np.uint64)
    with np.errstate(over="ignore"):
        m = (np.arange(1, v.size + 1, dtype=np.uint64)) * np.uint64(CHECKSUM_K)
        return int(((v + np.uint64(1)) * m).sum(dtype=np.uint64))


class McrawError(RuntimeError):
    pass


class Context:
    """One decoder context on one CUDA device (mcraw_ctx)."""

    def __init__(self, device=0):
        self._c = lib()
        h = ctypes.c_void_p()
        rc = self._c.mcraw_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise McrawError(f"mcraw_ctx_create({device}) failed ({rc}): {self._c.mcraw_last_error(None).decode()}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._c.mcraw_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise McrawError(f"{what} failed ({rc}): {self._c.mcraw_last_error(self._h).decode()}")

    # -- memory -------------------------------------------------------------------------------------
    def device_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self._c.mcraw_device_alloc(self._h, nbytes, ctypes.byref(p)), "mcraw_device_alloc")
        return p.value

    def device_free(self, ptr):
        self._check(self._c.mcraw_device_free(self._h, ptr), "mcraw_device_free")

    def pinned_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self._c.mcraw_host_alloc_pinned(self._h, nbytes, ctypes.byref(p)), "mcraw_host_alloc_pinned")
        return p.value

    def pinned_free(self, ptr):
        self._check(self._c.mcraw_host_free_pinned(self._h, ptr), "mcraw_host_free_pinned")

    def pinned_array(self, nbytes):
        """A pinned host buffer viewed as a numpy uint8 array (freed with the context owner's pinned_free)."""
        ptr = self.pinned_alloc(nbytes)
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=np.uint8)
        return ptr, arr

    def h2d(self, dst_dev, src, stream=None, sync=True):
        src = np.ascontiguousarray(src)
        self._check(self._c.mcraw_memcpy_h2d(self._h, dst_dev, src.ctypes.data, src.nbytes, stream), "mcraw_memcpy_h2d")
        if sync:
            self.sync(stream)

    def d2h(self, dst, src_dev, nbytes=None, stream=None, sync=True):
        assert dst.flags["C_CONTIGUOUS"]
        self._check(self._c.mcraw_memcpy_d2h(self._h, dst.ctypes.data, src_dev, dst.nbytes if nbytes is None else nbytes,
                                             stream), "mcraw_memcpy_d2h")
        if sync:
            self.sync(stream)

    def sync(self, stream=None):
        self._check(self._c.mcraw_stream_sync(self._h, stream), "mcraw_stream_sync")

    # -- decode -------------------------------------------------------------------------------------
    @staticmethod
    def make_descs(frames):
        """frames: iterable of (src_ptr, len, width, height, compression_type, dst_ptr, dst_capacity_elems[, encoded_width])."""
        frames = list(frames)
        arr = (FrameDesc * max(1, len(frames)))()
        for i, fr in enumerate(frames):
            src, ln, w, h, ct, dst, cap = fr[:7]
            arr[i].src, arr[i].len, arr[i].width, arr[i].height = src, ln, w, h
            arr[i].compression_type, arr[i].dst, arr[i].dst_capacity_elems = ct, dst, cap
            arr[i].encoded_width = fr[7] if len(fr) > 7 else 0
        return arr, len(frames)

    def decode_batch(self, descs, n, stream=None):
        self._check(self._c.mcraw_decode_batch(self._h, descs, n, stream), "mcraw_decode_batch")

    def decode_batch_levels(self, descs, levels, n, stream=None):
        """mcraw_decode_batch_levels; levels: list of (black[4], white, mode) per frame."""
        arr = (Levels * max(1, n))()
        for i, (black, white, mode) in enumerate(levels):
            arr[i].black[:] = [float(v) for v in black]
            arr[i].white, arr[i].mode = float(white), int(mode)
        self._check(self._c.mcraw_decode_batch_levels(self._h, descs, arr, n, stream), "mcraw_decode_batch_levels")

    def decode_batch_host(self, descs, n, stream=None):
        self._check(self._c.mcraw_decode_batch_host(self._h, descs, n, stream), "mcraw_decode_batch_host")

    def decode_batch_host_out(self, descs, host_dst_ptrs, n, stream=None):
        """Host in, host out (mcraw_decode_batch_host_out); host_dst_ptrs: ctypes array of n host pointers."""
        self._check(self._c.mcraw_decode_batch_host_out(self._h, descs, host_dst_ptrs, n, stream), "mcraw_decode_batch_host_out")

    def batch_wait(self, n):
        written = (ctypes.c_uint64 * max(1, n))()
        status = (ctypes.c_uint32 * max(1, n))()
        self._check(self._c.mcraw_batch_wait(self._h, written, status, n), "mcraw_batch_wait")
        return list(written[:n]), list(status[:n])

    def checksum_frames(self, ptrs, elems, stream=None):
        """mcraw_checksum_frames over device frames -> list of 64-bit sums (see checksum_u16)."""
        n = len(ptrs)
        pa = (ctypes.c_void_p * max(1, n))(*ptrs)
        ea = (ctypes.c_uint64 * max(1, n))(*elems)
        out = (ctypes.c_uint64 * max(1, n))()
        self._check(self._c.mcraw_checksum_frames(self._h, pa, ea, n, out, stream), "mcraw_checksum_frames")
        return list(out[:n])

    def decode_host(self, stream_bytes, width, height, compression_type, fill=0xA5A5):
        """Reference-shaped call: host bytes in, host uint16 image out -> (elements_written, image)."""
        src = np.ascontiguousarray(stream_bytes, dtype=np.uint8)
        out = np.full(width * height + 64, fill, dtype=np.uint16)
        n = self._c.mcraw_decode_host(self._h, out.ctypes.data, width, height, src.ctypes.data, src.size, compression_type)
        assert np.all(out[width * height:] == fill)
        return int(n), out[:width * height].reshape(height, width)

    @property
    def kernel_launches(self):
        return int(self._c.mcraw_kernel_launches(self._h))

    def kernel_time_totals(self):
        """(metadata-kernel ms, main-kernel ms, chunks) accumulated since the context was created."""
        a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64()
        self._check(self._c.mcraw_kernel_time_totals(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n)),
                    "mcraw_kernel_time_totals")
        return a.value, b.value, n.value

    def set_kernel_timing(self, every_n_chunks):
        """Bracket the kernels of every n-th chunk with CUDA events (0 = off, the default)."""
        self._check(self._c.mcraw_set_kernel_timing(self._h, every_n_chunks), "mcraw_set_kernel_timing")

    @property
    def last_batch_kernel_ms(self):
        return float(self._c.mcraw_last_batch_kernel_ms(self._h))


class DeviceBatch:
    """Test/bench helper: uploads compressed frames to device memory and owns the output buffers."""

    def __init__(self, ctx, frames):
        """frames: list of (np.uint8 stream, width, height, compression_type)."""
        self.ctx = ctx
        self.frames = frames
        self.src_ptrs, self.dst_ptrs = [], []
        items = []
        for stream, w, h, ct, *rest in frames:
            stream = np.ascontiguousarray(stream, dtype=np.uint8)
            sp = ctx.device_alloc(stream.size + 16)
            dp = ctx.device_alloc(w * h * 2 + 16)
            ctx.h2d(sp, stream)
            self.src_ptrs.append(sp)
            self.dst_ptrs.append(dp)
            # the uploader has seen the frame header: it tells the decoder which encodedWidth to plan for
            ew = rest[0] if rest else (frame_encoded_width(stream, w, h) if ct == COMPRESSION_CURRENT else 0)
            items.append((sp, stream.size, w, h, ct, dp, w * h, ew))
        self.descs, self.n = Context.make_descs(items)

    def decode(self, stream=None):
        self.ctx.decode_batch(self.descs, self.n, stream)
        return self.ctx.batch_wait(self.n)

    def fetch(self, i, fill=None):
        w, h = self.frames[i][1], self.frames[i][2]
        out = np.empty(w * h, dtype=np.uint16)
        self.ctx.d2h(out, self.dst_ptrs[i])
        return out.reshape(h, w)

    def fill_outputs(self, value=0xA5A5):
        for fr, dp in zip(self.frames, self.dst_ptrs):
            self.ctx.h2d(dp, np.full(fr[1] * fr[2], value, dtype=np.uint16))

    def free(self):
        for p in self.src_ptrs + self.dst_ptrs:
            self.ctx.device_free(p)
        self.src_ptrs, self.dst_ptrs = [], []
