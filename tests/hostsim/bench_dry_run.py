#!/usr/bin/env python
"""TEST INFRASTRUCTURE: a dry run of bench.py in a container without a GPU.  The library is the CPU build of the CUDA sources
(tests/hostsim/build.py), torch.cuda is replaced by stand-ins (host memory, wall-clock events).  Every number it prints is
meaningless; the point is that every code path of bench.py executes and the JSON line is well-formed before the round's only
real run on the GPU box.

  python tests/hostsim/bench_dry_run.py [bench.py flags; default: --log2-beliefs 11 --steps 2 --warmup 3]
"""
import contextlib
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import build as hostsim_build  # noqa: E402

so = hostsim_build.build()
os.environ["CCK_SO"] = so
os.environ["CCK_HOSTSIM"] = "1"
os.environ.setdefault("CCK_HOSTSIM_DEVICES", os.environ.get("WORLD_SIZE", "1"))
os.environ["LD_LIBRARY_PATH"] = os.path.dirname(so) + ":" + os.environ.get("LD_LIBRARY_PATH", "")

import torch  # noqa: E402


class FakeStream:
    cuda_stream = 1

    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass

    def record_event(self, e=None):
        e = e or FakeEvent()
        e.record()
        return e


class FakeEvent:
    def __init__(self, *a, **k):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return max(1e-3, (other.t - self.t) * 1e3)


_cpu = torch.device("cpu")
torch.device = lambda *a, **k: _cpu
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.Stream = FakeStream
torch.cuda.Event = FakeEvent
torch.cuda.set_stream = lambda s: None
torch.cuda.current_stream = lambda *a, **k: FakeStream()
torch.cuda.stream = lambda s: contextlib.nullcontext()
torch.cuda.get_device_name = lambda *a, **k: "hostsim"
torch.cuda.device_count = lambda: 1
torch.Tensor.pin_memory = lambda self, *a, **k: self
# multi-rank dry run (python -m torch.distributed.run --nproc-per-node 2 tests/hostsim/bench_dry_run.py --gpus 2 ...): gloo
# stands in for NCCL, the collectives and the rank logic of bench.py are the real ones
import torch.distributed as _dist  # noqa: E402

_init = _dist.init_process_group
_dist.init_process_group = lambda backend=None, **k: _init("gloo", **{kk: v for kk, v in k.items() if kk != "device_id"})

if len(sys.argv) == 1:
    sys.argv += ["--log2-beliefs", "11", "--steps", "2", "--warmup", "3"]
sys.argv[0] = os.path.join(ROOT, "bench.py")
import runpy  # noqa: E402

runpy.run_path(sys.argv[0], run_name="__main__")
