#!/bin/bash
# ncu evidence: (1) launch list of the bench command, (2) full-set captures of the conv kernels.
set +e
TAG=${1:-run}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_ncu.log 2>&1
echo "launch list exit=$?"
python tools/conv_cases.py --reps 2 > gpurun_out/${TAG}_cases_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|conv_direct_kernel|conv_dw_kernel' -c 40 \
    -o gpurun_out/${TAG}_conv python tools/conv_cases.py --reps 2 > gpurun_out/${TAG}_cases_ncu.log 2>&1
echo "full set exit=$?"
cat gpurun_out/${TAG}_cases_plain.log
tail -3 gpurun_out/${TAG}_cases_ncu.log
