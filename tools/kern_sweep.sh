#!/bin/bash
# GPU box: time the linear+Blackmore rollout kernel variants (CCK_KERNEL=3 state-in-smem, 2 grid, 1 group filter, 0 generic)
for k in ${KERNS:-2 1}; do
  for v in all_survive realistic; do
    CCK_KERNEL=$k timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --log2-beliefs ${LOG2B:-22} --no-e2e --no-cpu --no-realistic --variant $v 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('KERNEL=$k', '$v', 'Gsteps/s=%.3f' % (d['value']/1e9), 'kernel_ms=%.2f' % d['kernel_ms'], d['belief_steps_per_launch'])"
  done
done
