for b in 32 48 64 128 256; do for m in default 0; do
  if [ $m = default ]; then unset UYD_C3K_TC; else export UYD_C3K_TC=0; fi
  python bench.py --batch $b --steps 10 --warmup 3 --int8 0 --custom 0 --stress 0 --c4-batch 0 --sustain 0 --cpu-sample 4 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b tc=$m', round(d['ms_per_step'],4), round(d['value']))"
done; done
