#!/bin/bash
# Round evidence on the GPU box: tests, smoke, layer diff, bench lines, ncu launch list + --set full captures.
# Usage: bash tools/gpu_evidence.sh <tag>      (one GPU; every ncu pass runs after its own command exited 0 without ncu)
set +e
TAG=${1:-ev}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/${TAG}_pytest.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 gpurun_out/${TAG}_smoke.log)"
python tools/layer_diff.py > gpurun_out/${TAG}_layerdiff.log 2>&1; echo "layerdiff exit=$? :: $(tail -1 gpurun_out/${TAG}_layerdiff.log)"
python bench.py --profile-out gpurun_out/${TAG}_ops.md > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.log 2>&1; echo "bench ref exit=$? :: $(tail -1 gpurun_out/${TAG}_bench_ref.log | cut -c1-200)"
python tools/int8_table.py > gpurun_out/${TAG}_int8_ops.txt 2>&1; echo "int8 table exit=$?"
python tools/kernel_table.py --top 40 > gpurun_out/${TAG}_ktable.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/kernel_table.py --ncu > gpurun_out/${TAG}_launches.log 2>&1; echo "launch list exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"stem_v2|nms_image|conv_chain" -c 10 \
    -o gpurun_out/${TAG}_a -f python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full_a.log 2>&1; echo "full set a exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"c3k_flat|c3k_tc|conv_tc_kernel|sppf|conv_dw" -c 16 \
    -o gpurun_out/${TAG}_b -f python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full_b.log 2>&1; echo "full set b exit=$?"
python tools/kernel_table.py --custom --top 20 > gpurun_out/${TAG}_ktable_custom.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_big|conv_stem_mma" -c 8 \
    -o gpurun_out/${TAG}_c -f python tools/kernel_table.py --custom --ncu > gpurun_out/${TAG}_full_c.log 2>&1; echo "full set c exit=$?"
[ -n "$UYD_EVIDENCE_FAST" ] || { tools/probes/mma_rate > gpurun_out/${TAG}_mma_rate.txt 2>&1; echo "mma_rate exit=$?"; }
[ -n "$UYD_EVIDENCE_FAST" ] || { bash tools/c3k_sweep.sh > gpurun_out/${TAG}_c3k_sweep.txt 2>&1; echo "c3k sweep exit=$?"; }
du -sh gpurun_out
