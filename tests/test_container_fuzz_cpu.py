"""CPU: mutated .mcraw containers through the drop-in Decoder's host side (open / index / audio / frame read + JSON):
every failure is an exception, never an abort, a fault or a runaway allocation (tests/helpers/container_fuzz.py)."""
import os
import subprocess
import sys

import pytest

from motioncam_decoder_b200 import _lib

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.LIB_DROPIN), reason="drop-in library not built (needs nvcc)")
HELPER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "container_fuzz.py")


@pytest.mark.parametrize("seed", [2, 3])
def test_mutated_containers_fail_cleanly(tmp_path, seed):
    r = subprocess.run([sys.executable, HELPER, str(seed), "500", str(tmp_path)], capture_output=True, text=True, timeout=600)
    last = r.stderr.strip().splitlines()[-3:]
    assert r.returncode == 0, f"child died (rc {r.returncode}) at mutant/kind {last}"
    assert r.stdout.startswith("ok "), r.stdout
