#!/bin/bash
# A/B of env-switched kernel variants on the GPU box: the per-op table of the bench step for each environment.
# Usage: bash tools/gpu_ab.sh <tag> "ENV1=1" "ENV2=1 ENV3=1" ...   ("-" = no switch)
set +e
TAG=$1; shift
mkdir -p gpurun_out
i=0
for E in "$@"; do
  [ "$E" = "-" ] && E=""
  env $E python bench.py --int8 0 --custom 0 --stress 0 --c4-batch 0 --sustain 0 --cpu-sample 0 --profile-out gpurun_out/${TAG}_$i.md > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
  echo "== [$i] '$E' exit=$? :: $(python -c "
import json
d = json.loads(open('gpurun_out/${TAG}_$i.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 4), ' '.join(f'{k}={v[\"ms\"]:.4f}' for k, v in d['roofline']['families'].items()))
")"
  i=$((i+1))
done
python - "$TAG" "$i" <<'PY'
import sys, re
tag, n = sys.argv[1], int(sys.argv[2])
tabs = []
for i in range(n):
    d = {}
    for l in open(f"gpurun_out/{tag}_{i}.md"):
        m = re.match(r"\| (\d+) \| (.*?) \| ([\d.]+) \|", l)
        if m: d[int(m.group(1))] = (m.group(2), float(m.group(3)))
    tabs.append(d)
for k in sorted(tabs[0], key=lambda k: -tabs[0][k][1]):
    print(f"{k:3d} {tabs[0][k][0][:58]:58s} " + " ".join(f"{t.get(k, ('', 0))[1]*1e3:7.1f}" for t in tabs))
PY
