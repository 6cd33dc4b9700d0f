#!/bin/bash
# GPU box: ncu launch list + one --set full capture of the unicycle rollout kernel.  tools/profile_uni.sh TAG
set -u
TAG=${1:-r1_uni}
CMD="python bench.py --model unicycle --steps 2 --warmup 3 --log2-beliefs 20 --no-e2e --no-cpu --no-realistic"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o gpurun_out/${TAG}_rollout $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
