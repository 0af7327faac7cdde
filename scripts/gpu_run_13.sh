set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --dump-launches gpurun_out/launches_eager.csv > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_step']):
    print(f"{k:16s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:5.0f}  {v['gbs']:8.1f} GB/s")
PY
