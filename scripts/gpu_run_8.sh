# session-8: tests, bench with per-launch dump, ncu launch list + dram-traffic pass + one --set full capture
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 > gpurun_out/t_all.log
grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/launches_eager.csv > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_step']):
    print(f"{k:16s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:5.0f}  {v['gbs']:8.1f} GB/s")
PY
# one step under ncu: duration + dram traffic per launch (a few passes per kernel)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --launch-skip 1000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_bench.out 2>&1
wc -l gpurun_out/launches.csv
# --set full capture of the top kernels (a handful of launches each)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"d2_bwd_data|d2_bwd_weight|gemm_nt_tc|gemm_tn_tc" --launch-skip 353 -c 16 -o gpurun_out/top_full python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_full.out 2>&1
ls -la gpurun_out/
