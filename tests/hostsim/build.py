#!/usr/bin/env python
"""TEST INFRASTRUCTURE: compile the product's CUDA sources for the CPU (tests/hostsim/cuda_sim.hpp) into
tests/hostsim/_build/libcck_b200.so, so that the C-ABI plumbing and the kernels' control flow can be exercised by the `-m gpu`
tests in a container without a GPU.  The only source transformation is mechanical:
  kernel<<<grid, block, smem, stream>>>(args)   ->  sim::launch(sim::cfg(grid, block, smem, stream), kernel, args)
  extern __shared__ T name[];                   ->  T* name = (T*)sim::dyn_smem();
Everything else is compiled as written (g++ -ffp-contract=off stands in for nvcc -fmad=false).

  python tests/hostsim/build.py            # build if stale
  CCK_SO=tests/hostsim/_build/libcck_b200.so CCK_HOSTSIM=1 python -m pytest tests -m gpu -q
"""
import glob
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "chance-constrained-k-cbs_b200", "csrc")
# CCK_HOSTSIM_ASAN=1: AddressSanitizer build (out-of-bounds accesses of "device" buffers abort with a report), run with
#   LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
ASAN = os.environ.get("CCK_HOSTSIM_ASAN", "") == "1"
OUT = os.path.join(HERE, "_build", "asan") if ASAN else os.path.join(HERE, "_build")
SO = os.path.join(OUT, "libcck_b200.so")

LAUNCH = re.compile(r"(\b[A-Za-z_]\w*(?:<[^<>();]*>)?)\s*<<<(.+?)>>>\s*\(")
DYN = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];")


def transform(text):
    text = LAUNCH.sub(lambda m: f"sim::launch(sim::cfg({m.group(2)}), {m.group(1)}, ", text)
    text = text.replace("__noinline__", "CCK_SIM_NOINLINE")
    text = DYN.sub(lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)sim::dyn_smem();", text)
    return text


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    deps = srcs + [os.path.join(HERE, "cuda_sim.hpp"), os.path.abspath(__file__), os.path.join(ROOT, "include", "cck_b200.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    # generated tree: _build/src/csrc/*, so that the sources' own "../../include/cck_b200.h" resolves to _build/include (a link)
    gen = os.path.join(OUT, "src", "csrc")
    os.makedirs(gen, exist_ok=True)
    for s in srcs:
        name = os.path.basename(s)
        if name.endswith(".cu"):
            name = name[:-3] + "_sim.cpp"
        open(os.path.join(gen, name), "w").write(transform(open(s).read()))
    link = os.path.join(OUT, "include")
    if not os.path.lexists(link):
        os.symlink(os.path.join(ROOT, "include"), link)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unused-function",
           "-Wno-unused-variable", "-Wno-unknown-pragmas", "-Wno-unused-but-set-variable", "-Wno-sign-compare",
           "-I", os.path.join(HERE, "shim"), "-include", os.path.join(HERE, "cuda_sim.hpp"),
           "-o", SO, os.path.join(gen, "cck_b200_sim.cpp"), "-pthread"]
    if ASAN:
        cmd[1:1] = ["-fsanitize=address", "-fno-omit-frame-pointer"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-6000:], out.stderr[-12000:])
    if out.returncode != 0:
        raise RuntimeError("hostsim build failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
