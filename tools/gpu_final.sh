#!/bin/bash
# Last pass of a round: full GPU tests, smoke, the bench line with the per-op tables, INT8 plan table, ncu of the kernels
# added in the last pass.  Usage: bash tools/gpu_final.sh <tag>
set +e
TAG=${1:-fin}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/${TAG}_pytest.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 gpurun_out/${TAG}_smoke.log)"
python bench.py --profile-out gpurun_out/${TAG}_ops.md > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>&1; echo "bench ref exit=$?"
python tools/int8_table.py 64 --all > gpurun_out/${TAG}_int8_ops.txt 2>&1; echo "int8 table exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"c3k_flat_q_kernel" -c 8 \
    -o gpurun_out/${TAG}_q -f python tools/int8_table.py 64 --ncu > gpurun_out/${TAG}_full_q.log 2>&1; echo "full set q exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_dw_col_kernel|sppf_plane" -c 2 \
    -o gpurun_out/${TAG}_d -f python tools/kernel_table.py --ncu > gpurun_out/${TAG}_full_d.log 2>&1; echo "full set d exit=$?"
du -sh gpurun_out
