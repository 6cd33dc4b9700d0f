#!/bin/bash
# GPU box: time the rollout kernel for each beliefs-per-thread setting (CCK_BPT=0 is the generic kernel)
for b in ${BPTS:-0 1 2 4}; do
  for v in all_survive realistic; do
    CCK_BPT=$b timeout 300 python bench.py --steps 3 --warmup 3 --log2-beliefs 22 --no-e2e --no-cpu --variant $v 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BPT=$b', '$v', 'Gsteps/s=%.3f' % (d['value']/1e9), 'kernel_ms=%.2f' % d['kernel_ms'], d['belief_steps_per_launch'])"
  done
done
