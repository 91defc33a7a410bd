#!/usr/bin/env python
"""mcraw_tool.py -- command-line front end of the B200 decoder, in the spirit of the reference's example.cpp
(/root/reference/example.cpp:141-203: open a file, list frames, dump the first N frames and the audio).

    mcraw_tool.py synth  out.mcraw [--frames 8] [--width 1920] [--height 1080] [--legacy] [--audio-chunks 4]
    mcraw_tool.py info   clip.mcraw
    mcraw_tool.py decode clip.mcraw [-n N] [--out DIR] [--batch 32]     (needs a B200: there is no CPU decode path)
    mcraw_tool.py export clip.mcraw [-n N] [--out DIR] [--batch 16] [--threads 4] [--no-audio]     (needs a B200)

`decode` writes frame_%06d.u16 (little-endian 16-bit Bayer, width*height) + frame_000000.json (the frame metadata) and
audio.wav.  `export` is the reference program itself: audio.wav + frame_%06d.dng, byte-identical to the files
example.cpp writes (motioncam::exportClip, include/motioncam/Export.hpp; the compiled form is mcraw_export).
"""
import argparse
import json
import os
import sys
import wave

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def cmd_synth(a):
    from motioncam_decoder_b200 import testvec as tv
    frames = []
    for i in range(a.frames):
        img = tv.gen_photon(a.width, a.height, a.maxval, seed=1 + i % 8)
        data = tv.encode_legacy(img) if a.legacy else tv.encode_current(img)
        frames.append({"timestamp": 1_000_000 + 33_333 * i, "data": data, "width": a.width, "height": a.height,
                       "compressionType": 6 if a.legacy else 7})
    rng = np.random.default_rng(1)
    audio = [(1_000_000 + 20_000 * k, rng.integers(-8000, 8000, 1920 * 2, dtype=np.int16)) for k in range(a.audio_chunks)]
    tv.write_mcraw(a.file, frames, audio)
    print(f"wrote {a.file}: {a.frames} frames {a.width}x{a.height} (compressionType {6 if a.legacy else 7}), "
          f"{a.audio_chunks} audio chunks, {os.path.getsize(a.file)} bytes")


def cmd_info(a):
    from motioncam_decoder_b200 import hostapi
    with hostapi.Decoder(a.file) as d:
        stamps = d.get_frames()
        meta = d.get_container_metadata()
        print(json.dumps({"frames": len(stamps), "first_timestamp": stamps[0] if stamps else None,
                          "last_timestamp": stamps[-1] if stamps else None,
                          "audio_sample_rate_hz": meta.get("extraData", {}).get("audioSampleRate"),
                          "audio_channels": meta.get("extraData", {}).get("audioChannels"),
                          "audio_chunks": len(d.load_audio()), "container_metadata_keys": sorted(meta)}, indent=1))


def cmd_decode(a):
    from motioncam_decoder_b200 import hostapi
    os.makedirs(a.out, exist_ok=True)
    with hostapi.Decoder(a.file) as d:
        stamps = d.get_frames()
        if a.n is not None:
            stamps = stamps[:a.n]
        done = 0
        for k in range(0, len(stamps), a.batch):
            part = stamps[k:k + a.batch]
            for ts, data in zip(part, d.load_frames(part)):            # one batched device decode per `batch` frames
                data.tofile(os.path.join(a.out, f"frame_{done:06d}.u16"))
                done += 1
        # frame metadata through the reference-shaped call for the first frame only (cheap sanity line for the user)
        if stamps:
            _, meta = d.load_frame(stamps[0])
            with open(os.path.join(a.out, "frame_000000.json"), "w") as f:
                json.dump(meta, f)
        chunks = d.load_audio()
        if chunks:
            with wave.open(os.path.join(a.out, "audio.wav"), "wb") as w:
                w.setnchannels(d.num_audio_channels())
                w.setsampwidth(2)
                w.setframerate(d.audio_sample_rate_hz())
                for _, samples in chunks:
                    w.writeframes(samples.tobytes())
        print(f"decoded {done} frames and {len(chunks)} audio chunks into {a.out}")


def cmd_export(a):
    from motioncam_decoder_b200 import hostapi
    os.makedirs(a.out, exist_ok=True)
    n = hostapi.export_clip(a.file, a.out, num_frames=-1 if a.n is None else a.n, batch=a.batch, writer_threads=a.threads,
                            audio=not a.no_audio)
    print(f"exported {n} frames into {a.out}")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("synth")
    s.add_argument("file")
    s.add_argument("--frames", type=int, default=8)
    s.add_argument("--width", type=int, default=1920)
    s.add_argument("--height", type=int, default=1080)
    s.add_argument("--maxval", type=int, default=4095)
    s.add_argument("--legacy", action="store_true")
    s.add_argument("--audio-chunks", type=int, default=4)
    s.set_defaults(fn=cmd_synth)
    s = sub.add_parser("info")
    s.add_argument("file")
    s.set_defaults(fn=cmd_info)
    s = sub.add_parser("decode")
    s.add_argument("file")
    s.add_argument("-n", type=int, default=None, help="number of frames (like the reference example's -n)")
    s.add_argument("--out", default="mcraw_out")
    s.add_argument("--batch", type=int, default=32)
    s.set_defaults(fn=cmd_decode)
    s = sub.add_parser("export")
    s.add_argument("file")
    s.add_argument("-n", type=int, default=None, help="number of frames (like the reference example's -n)")
    s.add_argument("--out", default="mcraw_out")
    s.add_argument("--batch", type=int, default=16)
    s.add_argument("--threads", type=int, default=4)
    s.add_argument("--no-audio", action="store_true")
    s.set_defaults(fn=cmd_export)
    a = ap.parse_args()
    a.fn(a)


if __name__ == "__main__":
    main()
