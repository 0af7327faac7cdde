# round-end re-baseline: all GPU tests, ncu launch list of one step (time + DRAM traffic) -> profiles/,
# bench (reads profiles/ncu_traffic.json for roofline.traffic), reference arm, --set full capture of the top kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 800 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile --skip-infer > gpurun_out/ncu_bench.out 2>&1
python scripts/ncu_step_summary.py gpurun_out/launches.csv gpurun_out/launches_summary.txt gpurun_out/ncu_traffic.json | head -60
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/launches_eager.csv > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'], d['step_roofline'], d['cpu_baseline'], d['infer'])
for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_step']):
    print(f"{k:16s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:5.0f}  {v['gbs']:8.1f} GB/s")
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cut -c1-500 gpurun_out/bench_ref.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"d2_bwd_data|d2_bwd_weight|ww_conv|ww_wgrad|d2_fwd" --launch-skip 114 -c 20 -o gpurun_out/dw_full python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile --skip-infer > gpurun_out/ncu_full.out 2>&1
ls -la gpurun_out/
