/*
 * mcraw_b200.h -- C-ABI of the B200-native MCRAW frame decoder (libmcraw_b200.so).
 *
 * The reference (mirsadm/motioncam-decoder) has no FFI layer: its boundary for this path is two C++ free
 * functions and one class.  This header is the plain-C surface underneath our drop-in versions of those,
 * and the batched device entry point the reference does not have:
 *
 *   reference interface (file:line under /root/reference)            entry point here
 *   ---------------------------------------------------------------  -----------------------------------------
 *   raw::Decode        lib/include/motioncam/RawData.hpp:25-30       mcraw_decode_host(..., MCRAW_COMPRESSION_CURRENT)
 *                      lib/RawData.cpp:528-612
 *   raw::DecodeLegacy  lib/include/motioncam/RawData.hpp:32-37       mcraw_decode_host(..., MCRAW_COMPRESSION_LEGACY)
 *                      lib/RawData_Legacy.cpp:445-495
 *   Decoder::loadFrame lib/Decoder.cpp:184-235 (one frame per call,  mcraw_decode_batch / mcraw_decode_batch_host
 *                      host vectors)                                 (many frames per call, device output)
 *
 * Conventions (same as the reference): results are counted in uint16 ELEMENTS written; 0 means the frame
 * failed (RawData.cpp:547-554, Decoder.cpp:225-230).  Nothing here throws; functions returning int give
 * MCRAW_OK or a negative error and leave text in mcraw_last_error().
 *
 * There is no CPU decode path behind this API: every decode runs the sm_100a kernels, and context creation
 * fails when no CUDA device is usable.
 *
 * Threading: a context is not thread-safe; contexts are independent (one per GPU / host thread).
 */
#ifndef MCRAW_B200_H
#define MCRAW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCRAW_COMPRESSION_LEGACY 6  /* Decoder.cpp:20 MOTIONCAM_COMPRESSION_TYPE_LEGACY */
#define MCRAW_COMPRESSION_CURRENT 7 /* Decoder.cpp:21 MOTIONCAM_COMPRESSION_TYPE */

#define MCRAW_OK 0
#define MCRAW_ERR_CUDA (-1)     /* a CUDA runtime call failed */
#define MCRAW_ERR_ARG (-2)      /* bad argument (null, misaligned, unsupported geometry) */
#define MCRAW_ERR_NO_DEVICE (-3)
#define MCRAW_ERR_STATE (-4)    /* e.g. results requested with no batch in flight */

/* Per-frame failure reasons reported through mcraw_batch_wait's status[] (0 = decoded). */
#define MCRAW_FRAME_OK 0u
#define MCRAW_FRAME_BAD_HEADER 1u      /* offsets > len, encodedWidth % 64, encodedWidth < width (RawData.cpp:547-554) */
#define MCRAW_FRAME_TRUNCATED 2u       /* a block or metadata block runs past len (reference: stale data, RawData.cpp:419) */
#define MCRAW_FRAME_BAD_BITS 4u        /* bits[] value > 16 (reference: out-of-bounds table read) */
#define MCRAW_FRAME_BAD_META_COUNT 8u  /* metadata count smaller than the number of blocks */
#define MCRAW_FRAME_GEOMETRY 16u       /* encodedHeight larger than the descriptor's height allows, dst too small (legacy), or the
                                        * header's encodedWidth is not what the descriptor planned for (see encoded_width) */
#define MCRAW_FRAME_BAD_TYPE 32u       /* compression_type not 6 or 7 (Decoder.cpp:233) */
#define MCRAW_FRAME_INTERNAL 64u       /* a device-side wait gave up (never expected; reported instead of hanging the GPU) */

typedef struct mcraw_ctx mcraw_ctx; /* opaque; owns scratch, streams, staging rings on one device */

/* One frame of a batch.  For mcraw_decode_batch src and dst are DEVICE pointers, 16-byte aligned.
 * For mcraw_decode_batch_host src is HOST memory (pinned for overlap; pageable works but serialises). */
typedef struct mcraw_frame_desc {
    const uint8_t* src;          /* compressed frame buffer (what Decoder.cpp:199-201 freads into mTmpBuffer) */
    uint64_t len;                /* its size in bytes */
    int32_t width;               /* frame JSON "width"  (Decoder.cpp:216) */
    int32_t height;              /* frame JSON "height" (Decoder.cpp:217) */
    int32_t compression_type;    /* frame JSON "compressionType": 7 or 6 (Decoder.cpp:218) */
    int32_t encoded_width;       /* current format only.  0 = the frame header's encodedWidth is width rounded up to 64 (what every
                                  * known encoder writes).  Otherwise: the encodedWidth the caller has read from bytes 0..3 of the
                                  * frame (RawData.cpp:500-524) -- any multiple of 64 >= width is a valid frame for the reference
                                  * (RawData.cpp:550-554), and the work list is planned from it.  The kernels check the header
                                  * against it; a mismatch fails the frame with MCRAW_FRAME_GEOMETRY.  The host-source entry
                                  * points (mcraw_decode_batch_host, mcraw_decode_host) read the header themselves when this is 0. */
    uint16_t* dst;               /* DEVICE output, width*height uint16, row-major (Decoder.cpp:221-222) */
    uint64_t dst_capacity_elems; /* capacity of dst in uint16 elements */
} mcraw_frame_desc;

/* ---- context ---------------------------------------------------------------------------------------- */
int mcraw_ctx_create(int device, mcraw_ctx** out);
void mcraw_ctx_destroy(mcraw_ctx* ctx);
const char* mcraw_last_error(const mcraw_ctx* ctx); /* ctx may be NULL: error of the last failed create */
int mcraw_ctx_device(const mcraw_ctx* ctx);

/* ---- batched device entry point -------------------------------------------------------------------- */
/* Enqueue the decode of n frames on `stream` (a cudaStream_t passed as void*; NULL = the context's own
 * stream).  Asynchronous: returns once the work is enqueued.  Frames may mix sizes and compression types.
 * The work is ordered after everything enqueued on `stream` before the call (so descs[i].src may be produced there).
 * Back-to-back calls on one stream overlap where that cannot be observed: when a call presents the descriptors of the
 * previous call again, or writes a disjoint range of output addresses, its index kernel is launched as a programmatic
 * dependent of the previous call's pixel kernel and resolves the metadata while those pixels still stream
 * (MCRAW_CHAIN=0 in the environment switches this off).  That holds for NEW descriptors as well: their plan (device-side
 * copies of the descriptors, work list) is uploaded on a stream of the context's own and the kernels wait for it
 * themselves, so `stream` carries nothing but the two kernels of the call (MCRAW_PLAN_SIDE=0: upload on `stream`). */
int mcraw_decode_batch(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream);

/* For callers that upload compressed frames themselves: the value to put into mcraw_frame_desc.encoded_width for a
 * compressionType 7 frame whose first bytes are at `frame` in HOST memory (0 when the header is the usual width rounded
 * up to 64, or is one no valid frame can carry -- the kernels then reject the frame by their own checks). */
int32_t mcraw_frame_encoded_width(const uint8_t* frame, uint64_t len, int32_t width, int32_t height);

/* Same, but descs[i].src are HOST buffers: the context copies them to device staging on its side streams in
 * chunks (double-buffered) so transfer overlaps decode, then decodes into descs[i].dst (device). */
int mcraw_decode_batch_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t n, void* stream);

/* ---- optional epilogue: black / white level -------------------------------------------------------- */
/* What the reference's consumer side does with a decoded frame starts from the container's blackLevel[4] / whiteLevel
 * (example.cpp:66-67, handed to the DNG at :89-91).  mcraw_decode_batch_levels fuses that first step into the pixel
 * kernels, where the samples are in registers anyway: levels[i] applies to descs[i]; black[] is indexed by the CFA
 * position (y & 1) * 2 + (x & 1).  The output keeps 16 bits per pixel, same layout, same `written` counts.
 *   MCRAW_OUT_RAW        the raw values (what the reference delivers; same as mcraw_decode_batch)
 *   MCRAW_OUT_BLACK_SUB  uint16: min(max(v - b, 0), w - b) with b = lrintf(black[c]), w = lrintf(white)  (exact integers)
 *   MCRAW_OUT_NORM_F16   IEEE half: clamp(((float)v - black[c]) * (1.0f / (white - black[c])), 0, 1), fp32 arithmetic,
 *                        one rounding to half (round to nearest even); white <= black[c] gives scale 0 */
#define MCRAW_OUT_RAW 0u
#define MCRAW_OUT_BLACK_SUB 1u
#define MCRAW_OUT_NORM_F16 2u
typedef struct mcraw_levels {
    float black[4];
    float white;
    uint32_t mode; /* MCRAW_OUT_* */
} mcraw_levels;
int mcraw_decode_batch_levels(mcraw_ctx* ctx, const mcraw_frame_desc* descs, const mcraw_levels* levels, uint32_t n, void* stream);

/* Host in, HOST OUT -- the reference's own contract for a decoded frame (Decoder.cpp:221-230: outData is a host vector),
 * batched: like mcraw_decode_batch_host, and the pixels of frame i (width*height uint16) are also copied to host_dst[i]
 * (pinned for overlap) as soon as the chunk that holds the frame has been decoded -- the device->host copy of chunk c runs
 * beside the host->device copy and the decode of chunk c+1.  descs[i].dst still names the DEVICE buffer the frame is
 * decoded into.  mcraw_batch_wait returns once the host buffers are complete. */
int mcraw_decode_batch_host_out(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint16_t* const* host_dst, uint32_t n, void* stream);

/* One logical batch enqueued in pieces, for feeders that produce frames while earlier ones already travel and decode
 * (Decoder::loadFramesToDevice reads chunk c+1 of the file while chunk c is on the bus): mcraw_batch_begin announces
 * n_total frames, every mcraw_batch_append_host call enqueues frames [first_index, first_index + count) (descs[0] is
 * frame first_index; host sources, as mcraw_decode_batch_host), mcraw_batch_wait then reports all n_total frames --
 * frames that were never appended come back as failed with MCRAW_FRAME_BAD_TYPE. */
int mcraw_batch_begin(mcraw_ctx* ctx, uint32_t n_total);
int mcraw_batch_append_host(mcraw_ctx* ctx, const mcraw_frame_desc* descs, uint32_t first_index, uint32_t count, void* stream);

/* Wait for the most recently enqueued batch and fetch its per-frame results: written_elems[i] is the number
 * of uint16 elements written for frame i (0 = failed), status[i] the MCRAW_FRAME_* bits.  Either may be NULL. */
int mcraw_batch_wait(mcraw_ctx* ctx, uint64_t* written_elems, uint32_t* status, uint32_t n);

/* ---- reference-shaped single-frame call (host in, host out, synchronous) ---------------------------- */
/* Same contract as motioncam::raw::Decode / DecodeLegacy: output holds width*height uint16; returns the
 * element count written or 0.  H2D, decode and D2H all happen inside the call. */
size_t mcraw_decode_host(mcraw_ctx* ctx, uint16_t* output, int width, int height, const uint8_t* input, size_t len,
                         int compression_type);

/* ---- memory helpers (so a C caller needs no CUDA headers) ------------------------------------------- */
int mcraw_device_alloc(mcraw_ctx* ctx, size_t bytes, void** out);
int mcraw_device_free(mcraw_ctx* ctx, void* p);
int mcraw_host_alloc_pinned(mcraw_ctx* ctx, size_t bytes, void** out);
int mcraw_host_free_pinned(mcraw_ctx* ctx, void* p);
/* Page-lock memory the caller already owns (e.g. a read-only mmap of the .mcraw file), so that mcraw_decode_batch_host
 * can copy frames to the device straight from it -- the "GPU-direct feed" of SURVEY.md section 8f-3, in place of the
 * page-cache -> pinned-ring copy in front of the reference's fread (Decoder.cpp:203-209).  read_only != 0 is required
 * for PROT_READ mappings.  Fails (MCRAW_ERR_CUDA, text in mcraw_last_error) where the platform cannot pin the range. */
int mcraw_host_register(mcraw_ctx* ctx, void* p, size_t bytes, int read_only);
int mcraw_host_unregister(mcraw_ctx* ctx, void* p);
int mcraw_memcpy_h2d(mcraw_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes, void* stream);
int mcraw_memcpy_d2h(mcraw_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes, void* stream);
int mcraw_stream_sync(mcraw_ctx* ctx, void* stream);

/* ---- integrity ---------------------------------------------------------------------------------------- */
/* Position-weighted 64-bit checksum of n decoded frames in DEVICE memory, computed on the device (one kernel over all
 * frames, synchronous):  out[f] = sum over i < elems[f] of (frame_f[i] + 1) * ((i + 1) * 0x9E3779B97F4A7C15)  mod 2^64.
 * The multiplier is odd, so any single changed sample changes the sum; numpy restates it in one line (capi.checksum_u16).
 * bench.py uses it to check every pixel of every timed batch against the source images without copying 1-25 GB back. */
int mcraw_checksum_frames(mcraw_ctx* ctx, const uint16_t* const* frames_dev, const uint64_t* elems, uint32_t n, uint64_t* out,
                          void* stream);

/* ---- introspection ---------------------------------------------------------------------------------- */
/* Kernels of this library launched through ctx since creation (bench.py reports the per-step delta). */
uint64_t mcraw_kernel_launches(const mcraw_ctx* ctx);
/* CUDA-event time of the decode kernels of the last waited batch, in milliseconds (0 if unavailable). */
float mcraw_last_batch_kernel_ms(const mcraw_ctx* ctx);
/* Kernel timing is off by default (the event records cost ~8 us per batch on the stream).  every_n_chunks = n > 0 brackets
 * the index kernels and the pixel kernels of every n-th chunk with CUDA events; 0 switches it off again. */
int mcraw_set_kernel_timing(mcraw_ctx* ctx, uint32_t every_n_chunks);
/* Sum of CUDA-event times (ms) of the index kernels and of the pixel kernels over the TIMED chunks this context has
 * completed, and their number.  Waits for in-flight work.  bench.py reads deltas. */
int mcraw_kernel_time_totals(mcraw_ctx* ctx, double* meta_ms, double* main_ms, uint64_t* chunks);
const char* mcraw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MCRAW_B200_H */
