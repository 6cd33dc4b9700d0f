"""Randomised parity fuzz on the hardware (the last GPU tests to run): tests/hostsim/fuzz_rollout.py in its bit-exact modes -
random worlds and half-plane tables, means within 1e-12 .. 1e-3 of a decision threshold, degenerate and non-finite beliefs,
ragged step counts, trajectory mode; plan checkers, batched expansion, distance / tree queries, witness batch.  The same
script ran for hours on the CPU build of the kernels (profiles/r2g_fuzz_cpu_build.txt); here a few dozen cases per mode run
on the B200 itself, where sqrt.approx, the device math library and real warp scheduling are in play."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
FUZZ = os.path.join(ROOT, "tests", "hostsim", "fuzz_rollout.py")


@pytest.mark.parametrize("mode,cases", [("", 40), ("--pvc", 150), ("--expand", 120), ("--tree", 25), ("--witness", 40)])
def test_fuzz_mode_is_bit_identical_to_the_oracle(mode, cases):
    cmd = [sys.executable, FUZZ, "--cases", str(cases), "--seed", "20261020"] + ([mode] if mode else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert out.returncode == 0 and "fuzz ok" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    assert " 0 mismatches" in out.stdout
