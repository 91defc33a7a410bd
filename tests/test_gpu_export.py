"""GPU: the reference's example program (clip -> audio.wav + frame_%06d.dng, /root/reference/example.cpp:141-203) on the
batched B200 decode -- motioncam::exportClip through the flat wrapper and the mcraw_export executable -- against
the compiled reference program run on the same file: same progress lines, every file byte-identical."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from motioncam_decoder_b200 import _lib, hostapi, testvec as tv

pytestmark = pytest.mark.gpu
REF_EXAMPLE = os.path.join(os.path.dirname(ol.REF_SO), "ref_example")
OUR_EXPORT = os.path.join(_lib.PKG_DIR, "mcraw_export")


def _clip(tmp_path, n=7):
    frames = []
    sizes = [(1928, 16), (640, 12), (328, 48)]
    for k in range(n):
        w, h = sizes[k % len(sizes)]
        legacy = k % 3 == 1
        img = tv.gen_photon(w, h, 4095 if k % 2 else 1023, seed=140 + k)
        frames.append({"timestamp": 7_000_000 + 33_333 * ((k * 3) % n), "width": w, "height": h,
                       "compressionType": 6 if legacy else 7,
                       "data": tv.encode_legacy(img) if legacy else tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=k),
                       "asShotNeutral": [0.5 + 0.01 * k, 1.0, 0.61]})
    rng = np.random.default_rng(3)
    audio = [(1000 * i, rng.integers(-3000, 3000, 1920, dtype=np.int16)) for i in range(5)]
    cm = dict(tv.DEFAULT_CONTAINER_METADATA, colorMatrix2=[0.8, -0.2, 0.004, -0.35, 1.1, 0.25, 0.0, 0.15, 0.65], whiteLevel=4095.0)
    path = str(tmp_path / "clip.mcraw")
    tv.write_mcraw(path, frames, audio, container_metadata=cm)
    return path, n


def _reference_run(tmp_path, clip, extra=()):
    d = tmp_path / "ref"
    d.mkdir(exist_ok=True)
    r = subprocess.run([REF_EXAMPLE, clip, *extra], cwd=d, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return d, r.stdout


def _same_files(a, b):
    names = sorted(os.listdir(a))
    assert names == sorted(os.listdir(b))
    for name in names:
        assert open(os.path.join(a, name), "rb").read() == open(os.path.join(b, name), "rb").read(), name
    return names


@pytest.mark.skipif(not (ol.have_ref() and os.path.exists(REF_EXAMPLE)), reason="reference example program not built")
def test_export_clip_equals_reference_program(tmp_path):
    clip, n = _clip(tmp_path)
    ref_dir, _ = _reference_run(tmp_path, clip)
    out = tmp_path / "ours"
    out.mkdir()
    assert hostapi.export_clip(clip, out, batch=3, writer_threads=2) == n          # 3 batches, the last one short
    names = _same_files(ref_dir, out)
    assert names == ["audio.wav"] + [f"frame_{i:06d}.dng" for i in range(n)]


@pytest.mark.skipif(not (ol.have_ref() and os.path.exists(REF_EXAMPLE)), reason="reference example program not built")
def test_export_executable_equals_reference_program(tmp_path):
    assert os.path.exists(OUR_EXPORT), "mcraw_export not built (make -C motioncam-decoder_b200/csrc dropin)"
    clip, n = _clip(tmp_path)
    ref_dir, ref_out = _reference_run(tmp_path, clip, ("-n", "4"))
    out = tmp_path / "ours"
    out.mkdir()
    r = subprocess.run([OUR_EXPORT, clip, "-n", "4"], cwd=out, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ref_out                                                     # "Found 7 frames" + 4 "Writing ..." lines
    assert len(_same_files(ref_dir, out)) == 5


def test_export_without_reference(tmp_path):
    """Self-contained (no compiled reference needed): pixels inside every DNG equal the source images."""
    clip, n = _clip(tmp_path)
    out = tmp_path / "o"
    out.mkdir()
    assert hostapi.export_clip(clip, out, num_frames=5, batch=2, writer_threads=3, audio=False) == 5
    assert sorted(os.listdir(out)) == [f"frame_{i:06d}.dng" for i in range(5)]
    with hostapi.Decoder(clip) as d:
        stamps = d.get_frames()
        for i in range(5):
            data, meta = d.load_frame(stamps[i])
            blob = open(out / f"frame_{i:06d}.dng", "rb").read()
            assert blob[8:8 + data.size] == data.tobytes()
            assert len(blob) > 8 + data.size + 300
    with pytest.raises(hostapi.DecoderError):
        hostapi.export_clip(str(tmp_path / "missing.mcraw"), out)


RELINKED = os.path.join(os.path.dirname(ol.REF_SO), "ref_example_relinked")


@pytest.mark.skipif(not (os.path.exists(REF_EXAMPLE) and os.path.exists(RELINKED)), reason="reference programs not built")
def test_relinked_reference_program(tmp_path):
    """the reference's own example.cpp + lib/Decoder.cpp, unchanged, linked against this repo's
    library instead of lib/RawData*.cpp (oracle/Makefile `relinked`) -- every frame goes through the mangled
    motioncam::raw::Decode / DecodeLegacy symbols onto the GPU; output files identical to the all-CPU reference program."""
    clip, n = _clip(tmp_path)
    ref_dir, ref_out = _reference_run(tmp_path, clip)
    out = tmp_path / "relinked"
    out.mkdir()
    r = subprocess.run([RELINKED, clip], cwd=out, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ref_out
    assert len(_same_files(ref_dir, out)) == n + 1
