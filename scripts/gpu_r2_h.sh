# round 2, call H (1 GPU): A/B of the fused activation backward in the dgrad GEMM epilogue
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
for rep in 1 2; do
TD3D_FUSE_DACT=1 timeout 300 python bench.py --steps 30 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench_dact1_$rep.json 2> gpurun_out/bench_dact1.err; echo "rc=$?"; bench_line gpurun_out/bench_dact1_$rep.json
TD3D_FUSE_DACT=0 timeout 300 python bench.py --steps 30 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench_dact0_$rep.json 2> gpurun_out/bench_dact0.err; echo "rc=$?"; bench_line gpurun_out/bench_dact0_$rep.json
done
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_tc.py tests/test_gpu_effnet.py -q 2>&1 | tail -15 | cut -c1-300
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; bench_line gpurun_out/bench_b0.json
