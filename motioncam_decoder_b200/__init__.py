"""Import alias for the package directory ``motioncam-decoder_b200/`` (a hyphen is not importable).

``import motioncam_decoder_b200`` resolves every submodule from ``../motioncam-decoder_b200``.
"""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "motioncam-decoder_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
