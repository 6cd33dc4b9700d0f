#!/bin/bash
# Everything that can be verified WITHOUT a GPU, in one go (about 20 minutes): build, CPU suite, the -m gpu suite on the CPU
# build of the kernels (plain, permuted thread schedules, AddressSanitizer), bench.py dry runs (N = 1, 2), a short fuzz.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
python -m pytest tests -x -q -m "not gpu"
python tests/hostsim/build.py
export CCK_SO=$PWD/tests/hostsim/_build/libcck_b200.so CCK_HOSTSIM=1 LD_LIBRARY_PATH=$PWD/tests/hostsim/_build
python -m pytest tests -x -q -m gpu
CCK_HOSTSIM_SCHED=reverse  python -m pytest tests -x -q -m gpu -k "not full_size and not host_cpp and not zz_gpu_fuzz"
CCK_HOSTSIM_SCHED=random:7 python -m pytest tests -x -q -m gpu -k "not full_size and not host_cpp and not zz_gpu_fuzz"
python -c "import __graft_entry__ as g; g.smoke()"
python tests/hostsim/bench_dry_run.py > /dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 \
  tests/hostsim/bench_dry_run.py --gpus 2 --log2-beliefs 10 --steps 2 --warmup 3 > /dev/null
for mode in "" --generic --pvc --expand --tree --witness; do python tests/hostsim/fuzz_rollout.py --minutes 1 --seed 1 $mode | tail -n 1; done
CCK_HOSTSIM_ASAN=1 python tests/hostsim/build.py
LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:alloc_dealloc_mismatch=0 \
  CCK_SO=$PWD/tests/hostsim/_build/asan/libcck_b200.so LD_LIBRARY_PATH=$PWD/tests/hostsim/_build/asan \
  python -m pytest tests -x -q -m gpu -k "not full_size and not host_cpp and not zz_gpu_fuzz"
echo "check_cpu: all green"
