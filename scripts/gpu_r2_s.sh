# round 2, call S (1 GPU): TN (weight-gradient) GEMM with taller stages
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
timeout 600 python scripts/gemm_bench_tn.py > gpurun_out/gemm_bench_tn.txt 2>&1; cat gpurun_out/gemm_bench_tn.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err | cut -c1-300; bench_line gpurun_out/bench.json
TD3D_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err; echo "nooverlap rc=$?"; bench_line gpurun_out/bench_nooverlap.json
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; bench_line gpurun_out/bench_b0.json
