# round 2, call N / O (1 GPU): 256-row tiles and lane-local statistic sums in the tcgen05 NT GEMM; affine2 with batched loads
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
timeout 600 python scripts/gemm_bench3.py > gpurun_out/gemm_bench3.txt 2>&1; cat gpurun_out/gemm_bench3.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err | cut -c1-300; bench_line gpurun_out/bench.json
TD3D_TC_MSUB=1 timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_msub1.json 2> gpurun_out/bench_msub1.err; echo "msub1 rc=$?"; bench_line gpurun_out/bench_msub1.json
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; tail -n 3 gpurun_out/bench_infer.err | cut -c1-300; bench_line gpurun_out/bench_infer.json
TD3D_TC_MSUB=1 timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_infer_msub1.json 2> gpurun_out/bench_infer_msub1.err; echo "infer msub1 rc=$?"; bench_line gpurun_out/bench_infer_msub1.json
