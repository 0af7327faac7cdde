# round 2, call W (1 GPU): NT GEMM with 4 epilogue groups
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
        for k, v in sorted(kk.items(), key=lambda kv: -kv[1][key])[:8]:
            print(f"  {k:16s} {v[key]:8.3f} ms  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
( time timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_model.py tests/test_gpu_infer.py -q -m gpu -x ) > gpurun_out/t_tc.log 2>&1; tail -n 4 gpurun_out/t_tc.log | cut -c1-300
timeout 600 python scripts/gemm_bench3.py > gpurun_out/gemm_bench3_g4.txt 2>&1; cat gpurun_out/gemm_bench3_g4.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_g4.err | cut -c1-300; bench_line gpurun_out/bench_g4.json
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_infer_g4.json 2> gpurun_out/bench_infer_g4.err; echo "infer rc=$?"; bench_line gpurun_out/bench_infer_g4.json
