#!/bin/bash
# GPU box: time alternative builds of the library (CCK_SO=path) on both variants of the sweep.
#   SOS="a.so b.so" LOG2B=24 tools/so_sweep.sh
for so in ${SOS}; do
  for v in all_survive realistic; do
    CCK_SO=$PWD/$so timeout 300 python bench.py --steps ${STEPS:-5} --warmup 3 --log2-beliefs ${LOG2B:-24} --no-e2e --no-cpu --no-realistic --no-solve --variant $v 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', '$v', 'Gsteps/s=%.3f' % (d['value']/1e9), 'kernel_ms=%.3f' % d['kernel_ms'], d['clocks'])"
  done
done
