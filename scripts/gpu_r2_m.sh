# round 2, call M (1 GPU): elementwise kernels with compile-time activation / SE placement, loads-in-flight A/B
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round(d.get('step_roofline', {}).get('frac', 0), 4))
    kk = d.get('kernel_kinds') or {}
    if kk:
        key = 'ms_per_step' if 'ms_per_step' in next(iter(kk.values())) else 'ms_per_chunk'
        for k, v in sorted(kk.items(), key=lambda kv: -kv[1][key])[:12]:
            print(f"  {k:16s} {v[key]:8.3f} ms  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
( time timeout 2400 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_all.log 2>&1; tail -n 8 gpurun_out/t_all.log | cut -c1-300
for u in 4 2; do
  TD3D_EW_U=$u timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench_u$u.json 2> gpurun_out/bench_u$u.err; echo "u$u rc=$?"; tail -n 3 gpurun_out/bench_u$u.err | cut -c1-300; bench_line gpurun_out/bench_u$u.json
  TD3D_EW_U=$u timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_b0_u$u.json 2> gpurun_out/bench_b0_u$u.err; echo "b0 u$u rc=$?"; bench_line gpurun_out/bench_b0_u$u.json
done
( TD3D_EW_U=2 timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_effnet.py -q -m gpu -x ) > gpurun_out/t_u2.log 2>&1; tail -n 4 gpurun_out/t_u2.log | cut -c1-300
