# round 2, call B (1 GPU): fused depthwise backward parity + A/B, large-batch export hunt
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -k "depthwise" 2>&1 | tail -15 > gpurun_out/t_dw.log; tail -n 6 gpurun_out/t_dw.log
timeout 900 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -15 > gpurun_out/t_model.log; tail -n 6 gpurun_out/t_model.log
timeout 900 python -m pytest tests/test_gpu_surface.py -q 2>&1 | tail -40 > gpurun_out/t_surface.log; tail -n 25 gpurun_out/t_surface.log | cut -c1-400
for pf in 6 0 12; do
  TD3D_DWC_PF=$pf timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --dump-launches gpurun_out/launches_pf$pf.csv > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err; echo "pf=$pf rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pf$pf.json').read().strip().splitlines()[-1])
print('pf=$pf', d['value'], d['ms_per_step'])
for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]:
    print(f"  {k:16s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:5.0f}  {v['gbs']:8.1f} GB/s")
PY
done
grep dw_bwd gpurun_out/launches_pf6.csv
timeout 900 python -m pytest tests/test_gpu_infer.py -q 2>&1 | tail -30 > gpurun_out/t_infer.log; tail -n 20 gpurun_out/t_infer.log | cut -c1-400
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; tail -n 3 gpurun_out/bench_infer.err | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_infer.json').read().strip().splitlines()[-1])
    print('infer', d['value'], d['ms_per_step'], d['e2e'], d['step_roofline'], d['cpu_baseline'])
    for k,v in sorted(d['kernel_kinds'].items(), key=lambda kv:-kv[1]['ms_per_chunk']):
        print(f"  {k:16s} {v['ms_per_chunk']:8.3f} ms  n={v['launches_per_chunk']:5.0f}  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print('infer parse failed', e)
PY
for b in 512 1024; do
  CUDA_LAUNCH_BLOCKING=1 timeout 600 python tests/export_bigb.py $b > gpurun_out/bigb_$b.log 2>&1; echo "bigb $b rc=$?"; tail -n 4 gpurun_out/bigb_$b.log | cut -c1-600
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tests/export_bigb.py 1024 > gpurun_out/bigb_sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -E "Invalid|ERROR SUMMARY|at .*\.cu|BIGB|error" gpurun_out/bigb_sanitizer.log | head -30 | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"dwc_bwd" --launch-skip 45 -c 15 -o gpurun_out/dwc_full python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile --skip-infer > gpurun_out/ncu_dwc.out 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | tail -20
