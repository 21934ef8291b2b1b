#!/bin/bash
# Routine GPU-box pass: parity tests, smoke, layer diff, a short bench with the per-op table.
# Usage (on the GPU box): bash tools/gpu_round.sh [tag]
set +e
TAG=${1:-run}
mkdir -p gpurun_out
S=gpurun_out/${TAG}_summary.txt
: > $S
run() {
  name=$1; shift
  timeout 900 "$@" > gpurun_out/${TAG}_$name.log 2>&1
  echo "$name exit=$? :: $(tail -1 gpurun_out/${TAG}_$name.log | cut -c1-300)" | tee -a $S
}
run pytest python -m pytest tests -q -m gpu -x
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run layerdiff python tools/layer_diff.py
run bench python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/${TAG}_ops.md
cat $S
