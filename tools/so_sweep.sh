#!/bin/bash
# GPU box: time alternative builds of the library (CCK_SO=path) on the three variants of the sweep.
#   SOS="a.so b.so" LOG2B=24 tools/so_sweep.sh      ("default" = the in-tree libcck_b200.so)
for so in ${SOS}; do
  if [ "$so" = default ]; then unset CCK_SO; else export CCK_SO=$PWD/$so; fi
  timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 --log2-beliefs ${LOG2B:-24} --no-e2e --no-cpu --no-extra --no-solve 2>&1 | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so', ' '.join('%s=%.1fG/%.3fms' % (k, v['value']/1e9, v['kernel_ms']) for k, v in d['variants'].items()), d['clocks']['sm_mhz'])"
done
