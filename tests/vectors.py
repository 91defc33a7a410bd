"""Seeded parity vectors shared by the CPU tests (oracle vs compiled reference) and the GPU tests
(CUDA path vs oracle).  Each entry: (name, stream np.uint8, width, height, expected image or None)."""
import numpy as np

from motioncam_decoder_b200 import testvec as tv


def current_vectors(small=True):
    out = []
    # every header value 0..16 forced on every block (incl. the 7 / 9 / 11..15 aliases), one width per tile
    for hb in range(17):
        w_needed = {7: 7, 9: 9}.get(hb, hb if hb <= 10 else 16)
        widths = [min(w_needed, k) for k in range(w_needed + 1)]  # residual widths 0..w all stored at header hb
        img = tv.gen_forced_widths(256, 8, widths, seed=100 + hb)
        out.append((f"force_hdr{hb}", tv.encode_current(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=hb), 256, 8, img))
    # minimal headers for all residual widths 0..16 mixed inside one frame
    img = tv.gen_forced_widths(640, 16, list(range(17)), seed=5)
    out.append(("all_widths_minimal", tv.encode_current(img), 640, 16, img))
    out.append(("all_widths_aliases", tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=9), 640, 16, img))
    # geometry sweep: widths not multiples of 64 / 32 / 8, crop path, tiny frames
    for (w, h, mx, seed) in [(64, 4, 1023, 1), (100, 8, 1023, 2), (1928, 16, 4095, 3), (4080, 8, 1023, 4),
                             (8, 4, 255, 5), (2, 4, 1023, 6), (62, 12, 4095, 7), (1000, 4, 16383, 8),
                             (331, 8, 1023, 9), (4096, 4, 65535, 10)]:
        img = tv.gen_photon(w, h, mx, seed=seed)
        out.append((f"photon_{w}x{h}", tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=seed), w, h, img))
    # uint16 wrap of residual + reference (RawData.cpp:582-592)
    img = tv.gen_uniform(512, 16, 0, 65535, seed=11)
    out.append(("uniform16_wrap", tv.encode_current(img, ref_wrap=True, seed=12), 512, 16, img))
    img = tv.gen_photon(704, 12, 1023, seed=13)
    out.append(("photon_wrap", tv.encode_current(img, policy=tv.POLICY_ALIASES, ref_wrap=True, seed=14), 704, 12, img))
    # flat + noise (0-bit vs 10-bit blocks)
    img = tv.gen_flatnoise(1024, 32, cell=64, seed=15)
    out.append(("flatnoise", tv.encode_current(img), 1024, 32, img))
    # directly randomised well-formed streams: random bits (all of 0..16), random refs (any u16), random payload
    rng = np.random.default_rng(2024)
    for k, (ew, eh, w) in enumerate([(64, 4, 64), (192, 8, 150), (1024, 16, 1000), (2048, 8, 2048)]):
        nb = ew * eh // 64
        bits = rng.integers(0, 17, nb).astype(np.uint16)
        refs = rng.integers(0, 65536, nb).astype(np.uint16)
        out.append((f"random_stream_{k}", tv.assemble_current(ew, eh, bits, refs, seed=50 + k), w, eh, None))
    # metadata streams at odd byte offsets (RawData.cpp:463-498 reads bytes: any offset is valid for the reference)
    img = tv.gen_photon(1000, 16, 4095, seed=31)
    base = tv.encode_current(img, policy=tv.POLICY_ALIASES, seed=31)
    for pb, pr in [(1, 0), (0, 1), (1, 1), (3, 2), (5, 7), (2, 6)]:
        out.append((f"meta_pad_{pb}_{pr}", tv.pad_meta_current(base, pb, pr), 1000, 16, img))
    img16 = tv.gen_uniform(512, 8, 0, 65535, seed=32)                     # 16-bit metadata blocks at odd addresses
    out.append(("meta_pad_wrap16_1_1", tv.pad_meta_current(tv.encode_current(img16, ref_wrap=True, seed=33), 1, 1), 512, 8, img16))
    rngp = np.random.default_rng(34)
    nbp = 1024 * 16 // 64
    out.append(("meta_pad_random_stream_3_1", tv.pad_meta_current(
        tv.assemble_current(1024, 16, rngp.integers(0, 17, nbp).astype(np.uint16), rngp.integers(0, 65536, nbp).astype(np.uint16), seed=35),
        3, 1), 1000, 16, None))
    # encodedWidth larger than width rounded up to 64 (RawData.cpp:550-554 accepts any multiple of 64 >= width;
    # the rows are cropped to width, :598-608): frames encoded wider than they are decoded
    for (ew, w, h, seed) in [(128, 50, 8, 41), (192, 100, 8, 42), (1024, 520, 12, 43), (4096, 1928, 4, 44), (256, 8, 16, 45)]:
        wide = tv.gen_photon(ew, h, 1023, seed=seed)
        out.append((f"wide_enc_{ew}_as_{w}", tv.encode_current(wide, policy=tv.POLICY_ALIASES, seed=seed), w, h,
                    np.ascontiguousarray(wide[:, :w])))
    if not small:
        img = tv.gen_photon(1920, 1080, 4095, seed=21)
        out.append(("photon_1080p", tv.encode_current(img), 1920, 1080, img))
        img = tv.gen_photon(4080, 3072, 1023, seed=1234)
        out.append(("photon_c1", tv.encode_current(img), 4080, 3072, img))
    return out


def legacy_vectors(small=True):
    out = []
    for hb in range(16):
        w_needed = hb if hb <= 10 else 16
        widths = list(range(w_needed + 1))
        img = tv.gen_forced_widths(256, 8, widths, seed=200 + hb)
        # the legacy block reference is 12 bits: keep values where min <= 4095 is representable
        out.append((f"legacy_force_nib{hb}", tv.encode_legacy(img, policy=tv.POLICY_FORCE, policy_arg=hb, seed=hb), 256, 8, img))
    for (w, h, mx, seed) in [(32, 1, 1023, 1), (100, 3, 1023, 2), (1928, 5, 4095, 3), (4000, 6, 1023, 4),
                             (2, 2, 255, 5), (31, 7, 4095, 6), (33, 2, 1023, 7), (640, 9, 65535, 8)]:
        img = tv.gen_photon(w, h, mx, seed=seed)
        out.append((f"legacy_photon_{w}x{h}", tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=seed), w, h, img))
    img = tv.gen_flatnoise(1024, 16, cell=64, seed=15)
    out.append(("legacy_flatnoise", tv.encode_legacy(img), 1024, 16, img))
    out.append(("legacy_flatnoise_trailer", tv.encode_legacy(img, trailer_records=3), 1024, 16, img))
    img = tv.gen_uniform(512, 8, 0, 65535, seed=11)
    out.append(("legacy_uniform16", tv.encode_legacy(img, policy=tv.POLICY_ALIASES, seed=3), 512, 8, img))
    rng = np.random.default_rng(77)
    for k, (w, h) in enumerate([(32, 2), (150, 4), (1000, 8)]):
        nb = h * ((w + 31) // 32) * 2
        nib = rng.integers(0, 16, nb).astype(np.uint8)
        refs = rng.integers(0, 4096, nb).astype(np.uint16)
        out.append((f"legacy_random_stream_{k}", tv.assemble_legacy(w, h, nib, refs, seed=60 + k), w, h, None))
    if not small:
        img = tv.gen_photon(4000, 3000, 1023, seed=99)
        out.append(("legacy_c4", tv.encode_legacy(img), 4000, 3000, img))
    return out
