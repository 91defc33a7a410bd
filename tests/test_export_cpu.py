"""CPU: the consumer side of the path -- DNG / WAV packaging of decoded frames
(include/motioncam/Export.hpp) against the compiled reference program (/root/reference/example.cpp built into
oracle/_ref: its writeDng / writeAudio functions through oracle/ref_shim.cpp, and the whole program as
oracle/_ref/ref_example).  Files must be byte-identical.  Pixels come from the checker here; the GPU version of the
whole-program comparison is tests/test_gpu_export.py."""
import json
import os
import struct
import subprocess
import wave

import numpy as np
import pytest

import oracle_lib as ol
from motioncam_decoder_b200 import _lib, hostapi, testvec as tv

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_DROPIN), reason="drop-in library not built (needs nvcc)")
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="compiled reference not present")
REF_EXAMPLE = os.path.join(os.path.dirname(ol.REF_SO), "ref_example")


def _ref_lib():
    return hostapi.library(ol.REF_SO, "mcref_")


def _floats(rng, n):
    """Matrix entries across the whole float range: plain values, integers, zeros, very small (denominator beyond
    32 bits), tiny (denominator capped at 2^127), huge (numerator beyond 32 bits)."""
    out = []
    for kind in rng.integers(0, 8, n):
        if kind == 0:
            v = rng.uniform(-2, 2)
        elif kind == 1:
            v = float(np.float32(rng.uniform(-1, 1)) * np.float32(2.0) ** int(rng.integers(-40, 40)))
        elif kind == 2:
            v = float(rng.integers(-5, 5))
        elif kind == 3:
            v = float(np.float32(rng.uniform(0, 1)) * np.float32(2.0) ** int(rng.integers(-140, -100)))
        elif kind == 4:
            v = rng.uniform(0, 1e-3)
        elif kind == 5:
            v = rng.uniform(1e6, 1e12)
        elif kind == 6:
            v = 0.0
        else:
            v = rng.uniform(0.2, 3)
        out.append(float(v))
    return out


def _container(rng, it):
    black = [int(x) for x in rng.integers(0, 1024, 4)]
    if it % 5 == 1:
        black = [64.5, 63.25, 10, 11.9]                       # json floats -> uint16 like nlohmann does
    if it % 7 == 2:
        black = [int(x) for x in rng.integers(0, 65536, 4)]
    return {"blackLevel": black,
            "whiteLevel": float(rng.choice([1023.0, 4095.0, 65535.0, 16383.7, -3.0, 1e12, 70000.0, 32768.0])),
            "sensorArrangment": str(rng.choice(["rggb", "bggr", "grbg", "gbrg"])),
            "colorMatrix1": _floats(rng, 9), "colorMatrix2": _floats(rng, 9),
            "forwardMatrix1": _floats(rng, 9), "forwardMatrix2": _floats(rng, 9 + it % 3),
            "extraData": {"audioSampleRate": 48000, "audioChannels": 2}}


@needs_ref
def test_dng_bytes_match_reference_writer(tmp_path):
    rng = np.random.default_rng(11)
    a, b = str(tmp_path / "ref.dng"), str(tmp_path / "ours.dng")
    for it in range(200):
        w, h = int(rng.integers(1, 48)) * 2, int(rng.integers(1, 24)) * 2
        px = rng.integers(0, 65536, (h, w), dtype=np.uint16)
        cm = _container(rng, it)
        fm = {"width": w, "height": h, "asShotNeutral": _floats(rng, 3), "iso": 100}
        hostapi.write_dng(a, px, fm, cm, lib=_ref_lib(), prefix="mcref_")
        hostapi.write_dng(b, px, fm, cm)
        ra, rb = open(a, "rb").read(), open(b, "rb").read()
        assert ra == rb, (it, cm, fm)


def _parse_tiff(blob):
    assert blob[:4] == b"II*\0"
    ifd, = struct.unpack_from("<I", blob, 4)
    n, = struct.unpack_from("<H", blob, ifd)
    size = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 10: 8}
    fmt = {1: "B", 2: "c", 3: "H", 4: "I"}
    tags = {}
    prev = -1
    for i in range(n):
        tag, typ, count, slot = struct.unpack_from("<HHI4s", blob, ifd + 2 + 12 * i)
        assert tag > prev, "IFD entries must be sorted and unique"
        prev = tag
        nbytes = size[typ] * count
        raw = slot[:nbytes] if nbytes <= 4 else blob[struct.unpack("<I", slot)[0]:][:nbytes]
        assert len(raw) == nbytes
        if typ in (5, 10):
            v = struct.unpack("<" + ("II" if typ == 5 else "ii") * count, raw)
            vals = [(v[2 * k], v[2 * k + 1]) for k in range(count)]
        elif typ == 2:
            vals = raw
        else:
            vals = list(struct.unpack("<" + fmt[typ] * count, raw))
        tags[tag] = vals
    assert struct.unpack_from("<I", blob, ifd + 2 + 12 * n)[0] == 0          # single IFD
    assert ifd + 2 + 12 * n + 4 == len(blob)
    return tags


def test_dng_is_a_valid_cfa_tiff(tmp_path):
    """Independent of the reference: parse the file back and check the fields a DNG reader needs."""
    w, h = 200, 64
    img = tv.gen_photon(w, h, 1023, seed=3)
    cm = dict(tv.DEFAULT_CONTAINER_METADATA)
    cm["colorMatrix1"] = [0.75, -0.25, 0.125, -0.5, 1.5, 0.0625, 0.0, 0.3, 1.0]
    cm["sensorArrangment"] = "gbrg"
    path = str(tmp_path / "f.dng")
    hostapi.write_dng(path, img, {"width": w, "height": h, "asShotNeutral": [0.5, 1.0, 0.625]}, cm)
    blob = open(path, "rb").read()
    t = _parse_tiff(blob)
    assert t[256] == [w] and t[257] == [h] and t[258] == [16] and t[259] == [1] and t[262] == [32803]
    assert t[277] == [1] and t[278] == [h] and t[284] == [1] and t[254] == [0]
    assert t[273] == [8] and t[279] == [2 * w * h]
    assert np.array_equal(np.frombuffer(blob, np.uint16, w * h, 8).reshape(h, w), img)
    assert t[33421] == [2, 2] and t[33422] == [1, 2, 0, 1] and t[50711] == [1]
    assert t[50706] == [1, 4, 0, 0] and t[50707] == [1, 1, 0, 0] and t[50708] == b"MotionCam\0"
    assert t[50713] == [2, 2] and t[50714] == [64, 64, 64, 64] and t[50717] == [1023]
    assert t[50778] == [21] and t[50779] == [17] and t[50829] == [0, 0, h, w]
    for (num, den), want in zip(t[50721], cm["colorMatrix1"]):
        assert den > 0 and np.float32(num) / np.float32(den) == np.float32(want)
        assert num % 2 == 1 or den == 1 or num == 0                          # reduced by common powers of two
    assert t[50728] == [(1, 2), (1, 1), (5, 8)]
    assert [x[0] / x[1] for x in t[50964]] == cm["forwardMatrix1"]


def test_dng_metadata_errors():
    img = np.zeros((4, 4), np.uint16)
    good_c = dict(tv.DEFAULT_CONTAINER_METADATA)
    good_f = {"width": 4, "height": 4, "asShotNeutral": [1.0, 1.0, 1.0]}
    for key in ("blackLevel", "whiteLevel", "sensorArrangment", "colorMatrix1", "forwardMatrix2"):
        c = {k: v for k, v in good_c.items() if k != key}
        with pytest.raises(hostapi.DecoderError, match="Invalid container metadata"):
            hostapi.write_dng("/dev/null", img, good_f, c)
    for key in ("width", "height", "asShotNeutral"):
        f = {k: v for k, v in good_f.items() if k != key}
        with pytest.raises(hostapi.DecoderError, match="Invalid frame metadata"):
            hostapi.write_dng("/dev/null", img, f, good_c)
    with pytest.raises(hostapi.DecoderError, match="Invalid sensor arrangement"):          # example.cpp:105-106
        hostapi.write_dng("/dev/null", img, good_f, dict(good_c, sensorArrangment="xtrans"))
    with pytest.raises(hostapi.DecoderError, match="needs 9 values"):
        hostapi.write_dng("/dev/null", img, good_f, dict(good_c, colorMatrix2=[1.0] * 8))
    with pytest.raises(hostapi.DecoderError, match="Failed to open"):
        hostapi.write_dng("/nonexistent-dir/x.dng", img, good_f, good_c)


@needs_ref
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_wav_bytes_match_reference_writer(tmp_path, channels):
    rng = np.random.default_rng(channels)
    for nchunks in (0, 1, 5):
        chunks = [rng.integers(-32768, 32768, int(rng.integers(1, 400)) * 2, dtype=np.int16) for _ in range(nchunks)]
        if channels == 1 and chunks:
            chunks[0] = chunks[0][:-1]                                   # odd length is fine for mono
        a, b = str(tmp_path / "ref.wav"), str(tmp_path / "ours.wav")
        hostapi.write_audio(a, 48000, channels, chunks, lib=_ref_lib(), prefix="mcref_")
        hostapi.write_audio(b, 48000, channels, chunks)
        assert open(a, "rb").read() == open(b, "rb").read(), (channels, nchunks)


def test_wav_reads_back(tmp_path):
    rng = np.random.default_rng(9)
    chunks = [rng.integers(-32768, 32768, 960, dtype=np.int16) for _ in range(700)]      # more parts than one writev takes
    chunks.append(np.array([1, 2, 3], np.int16))                         # stereo: the unpaired last sample is dropped
    path = str(tmp_path / "a.wav")
    hostapi.write_audio(path, 44100, 2, chunks)
    with wave.open(path, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (2, 2, 44100)
        got = np.frombuffer(w.readframes(w.getnframes()), np.int16)
    want = np.concatenate(chunks[:-1] + [chunks[-1][:2]])
    assert np.array_equal(got, want)


@needs_ref
@pytest.mark.skipif(not os.path.exists(REF_EXAMPLE), reason="reference example program not built")
@pytest.mark.parametrize("legacy", [False, True])
def test_reference_program_files_equal_our_writers(tmp_path, legacy):
    """The reference's whole program (decode on the CPU + its DNG/WAV writers) on a synthetic clip, against this
    repo's container reader + writers fed with checker-decoded pixels: every output file byte-identical."""
    w, h, n = 328, 48, 3
    frames, images = [], []
    for i in range(n):
        img = tv.gen_photon(w, h, 4095, seed=70 + i)
        images.append(img)
        frames.append({"timestamp": 9_000 + 10 * (n - i), "width": w, "height": h, "compressionType": 6 if legacy else 7,
                       "data": tv.encode_legacy(img) if legacy else tv.encode_current(img),
                       "asShotNeutral": [0.51 + 0.01 * i, 1.0, 0.63]})
    rng = np.random.default_rng(2)
    audio = [(100 * k, rng.integers(-20000, 20000, 1920, dtype=np.int16)) for k in range(3)]
    cm = dict(tv.DEFAULT_CONTAINER_METADATA, colorMatrix1=[0.9, -0.3, 0.01, -0.4, 1.2, 0.2, 0.003, 0.1, 0.7], whiteLevel=4095.0)
    clip = str(tmp_path / "clip.mcraw")
    tv.write_mcraw(clip, frames, audio, container_metadata=cm)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    r = subprocess.run([REF_EXAMPLE, clip], cwd=ref_dir, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == [f"Found {n} frames"] + [f"Writing frame_{i:06d}.dng" for i in range(n)]

    with hostapi.Decoder(clip) as d:                                    # container side only: no frame decode on the CPU
        stamps = d.get_frames()
        container = d.get_container_metadata()
        chunks = [s for _, s in d.load_audio()]
        hostapi.write_audio(str(tmp_path / "audio.wav"), d.audio_sample_rate_hz(), d.num_audio_channels(), chunks)
    assert open(tmp_path / "audio.wav", "rb").read() == open(ref_dir / "audio.wav", "rb").read()
    by_ts = {f["timestamp"]: (f, img) for f, img in zip(frames, images)}
    for i, ts in enumerate(stamps):
        f, img = by_ts[ts]
        decode = ol.oracle_decode_legacy if legacy else ol.oracle_decode
        cnt, px = decode(f["data"], w, h)
        assert cnt == w * h and np.array_equal(px, img)
        fm = {"width": w, "height": h, "asShotNeutral": f["asShotNeutral"]}
        out = str(tmp_path / f"frame_{i:06d}.dng")
        hostapi.write_dng(out, px, fm, container)
        assert open(out, "rb").read() == open(ref_dir / f"frame_{i:06d}.dng", "rb").read(), i
    assert json.loads(json.dumps(container)) == cm


# ---- committed fixtures (tests/golden/export_golden.json, made by tests/golden/make_export_golden.py from the compiled
# ---- reference program): the pin that holds where neither /root/reference nor oracle/_ref exists
def _golden():
    import base64
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "export_golden.json")) as f:
        g = json.load(f)
    return g, base64.b64decode


@pytest.mark.parametrize("prefix", ["mcb200_", "mcref_"])
def test_golden_dng_and_wav(tmp_path, prefix):
    if prefix == "mcref_" and not ol.have_ref():
        pytest.skip("compiled reference not present")
    lib = _ref_lib() if prefix == "mcref_" else None
    g, dec = _golden()
    assert len(g["dng"]) == 6 and len(g["wav"]) == 5
    for c in g["dng"]:
        px = np.frombuffer(dec(c["pixels"]), np.uint16).reshape(c["frame"]["height"], c["frame"]["width"])
        path = str(tmp_path / (c["name"] + ".dng"))
        hostapi.write_dng(path, px, c["frame"], c["container"], lib=lib, prefix=prefix)
        assert open(path, "rb").read() == dec(c["file"]), c["name"]
    for c in g["wav"]:
        chunks = [np.frombuffer(dec(x), np.int16) for x in c["chunks"]]
        path = str(tmp_path / (c["name"] + ".wav"))
        hostapi.write_audio(path, c["rate"], c["channels"], chunks, lib=lib, prefix=prefix)
        assert open(path, "rb").read() == dec(c["file"]), c["name"]
