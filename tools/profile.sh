#!/bin/bash
# Run ON THE GPU BOX (via gpurun): plain run first, then the ncu launch list and one
# --set full capture of the rollout kernel.  Outputs land in gpurun_out/.
#   tools/profile.sh TAG [VARIANT] [LOG2B]
set -u
TAG=${1:-r1}
VARIANT=${2:-all_survive}
LOG2B=${3:-22}
CMD="python bench.py --steps 2 --warmup 3 --log2-beliefs $LOG2B --no-e2e --no-cpu --variant $VARIANT"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 2 -f -o gpurun_out/${TAG}_rollout $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | grep ${TAG}
