# session-6 re-baseline: GPU tests, bench, ncu launch list
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t_all.log
grep -E "passed|failed" gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_step']):
    print(f"{k:16s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:5.0f}  {v['gbs']:8.1f} GB/s")
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_bench.out 2>&1
wc -l gpurun_out/launches.csv
