"""The `-m gpu` parity tests, run HERE against a CPU build of the product's own CUDA sources (tests/hostsim: every CUDA thread
is a fiber, warp intrinsics and __syncthreads are rendez-vous points; the only source rewrite is the <<<...>>> launch syntax).

What this does and does not show: the C-ABI plumbing (table packing, struct layouts, argument order), the kernels' control flow
(work-counter refill, warp-uniform fast section, clearance field, broad phase, constraint / plan checkers) and their arithmetic as
g++ -ffp-contract=off evaluates it agree with the oracle and the committed reference vectors.  It is NOT the parity gate - that is
the same tests on a B200 - and it measures nothing.  It exists so that a container without a GPU still exercises the kernels after
every change (round 2 ran out of GPU minutes before the last kernels could be re-run on hardware)."""
import os
import re
import subprocess
import sys

from conftest import ROOT

HOSTSIM = os.path.join(ROOT, "tests", "hostsim")


def test_gpu_parity_tests_pass_on_the_cpu_build_of_the_cuda_sources():
    sys.path.insert(0, HOSTSIM)
    import build as hostsim_build
    so = hostsim_build.build()
    env = dict(os.environ, CCK_SO=so, CCK_HOSTSIM="1", LD_LIBRARY_PATH=os.path.dirname(so) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    # the full-size property sweep (2^24 beliefs), the C++ planner runs and the fuzz cases take minutes on fibers: run them by
    # hand (tests/hostsim/README.md; the fuzzer's long runs are recorded in profiles/r2g_fuzz_cpu_build.txt)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-q", "-x", "-m", "gpu", "-k", "not full_size and not host_cpp and not zz_gpu_fuzz",
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 80, tail
    assert "failed" not in out.stdout.splitlines()[-1], tail


def test_product_library_is_not_the_cpu_build():
    """The CPU build must never stand in for the product: it lives outside the package, is not shipped to the GPU box
    (.gpurunignore) and is only ever selected by an explicit CCK_SO."""
    assert "tests/hostsim/_build/" in open(os.path.join(ROOT, ".gpurunignore")).read()
    assert os.environ.get("CCK_SO", "") == "" or "hostsim" not in os.environ["CCK_SO"] or os.environ.get("CCK_HOSTSIM") == "1"
    import cck_b200
    if os.environ.get("CCK_SO", "") == "":
        assert cck_b200.capi._SO == os.path.join(ROOT, "chance-constrained-k-cbs_b200", "libcck_b200.so")
