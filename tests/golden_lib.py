"""Loader of the committed golden fixtures (tests/golden/mcraw_golden.npz, made by tests/golden/make_golden.py
from the compiled, unmodified reference)."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcraw_golden.npz")


def load():
    """-> list of (name, stream, width, height, compression_type, expected image)."""
    z = np.load(PATH)
    out = []
    for name in z["names"]:
        name = str(name)
        w, h, ct = (int(v) for v in z[name + ".meta"])
        out.append((name, z[name + ".stream"], w, h, ct, z[name + ".expect"]))
    return out
