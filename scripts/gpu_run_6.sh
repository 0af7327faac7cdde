# round-1 re-baseline: GPU tests, bench (both arms), ncu launch list, ncu --set full of the top kernels
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/t_all.log
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
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_bench.out 2>&1
wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"d2_bwd|d2_fwd|gemm_nt_tc|gemm_tn_tc|act_bwd_stats|affine2|apply_xform|stem_wgrad" -s 400 -c 60 -o gpurun_out/r01_top_full -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_full.out 2>&1
tail -n 3 gpurun_out/ncu_full.out | cut -c1-300
ls -la gpurun_out
