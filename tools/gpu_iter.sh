#!/bin/bash
# Quick iteration on the GPU box: the kernel parity tests, then the bench line with the per-op table.
# Usage: bash tools/gpu_iter.sh <tag> [pytest args]
set +e
TAG=${1:-it}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -x -m gpu "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/${TAG}_pytest.log)"
timeout 600 python bench.py --profile-out gpurun_out/${TAG}_ops.md > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "int8", round(d["int8"]["value"]), "custom", round(d["custom_variant"]["value"]), "c4", round(d["c4_batch256"]["value"]))
for k, v in d["roofline"]["families"].items():
    print(" ", k, v.get("launches"), round(v["ms"], 4))
PY
head -30 gpurun_out/${TAG}_ops.md
