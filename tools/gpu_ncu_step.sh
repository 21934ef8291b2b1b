#!/bin/bash
# ncu evidence for ONE predict step at the bench configuration (batch 64, 640x640):
#   (1) launch list (gpu__time_duration.sum) of every kernel of the step -> gpurun_out/<tag>_launches.csv
#   (2) --set full capture of the kernels matching <kernel-regex> (at most <count> launches; ~8 s
#       and ~1.6 MB each -- gpurun_out/ is limited to 64 MiB)             -> gpurun_out/<tag>_step.ncu-rep
# Usage (GPU box): bash tools/gpu_ncu_step.sh <tag> <kernel-regex> <count> [skip-launch-list]
set +e
TAG=${1:-run}
RE=${2:-conv_chain_kernel}
CNT=${3:-8}
mkdir -p gpurun_out
python tools/kernel_table.py --top 40 > gpurun_out/${TAG}_ktable.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_ktable.log; exit 1; }
if [ -z "$4" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/kernel_table.py --ncu > gpurun_out/${TAG}_launches.log 2>&1
echo "launch list exit=$?"
fi
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$RE" -c $CNT -o gpurun_out/${TAG}_step -f \
    python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full.log 2>&1
echo "full set exit=$?"
du -sh gpurun_out
