# round 2, call K (1 GPU): per-CTA statistic accumulation in the NT GEMM, depthwise forward dispatch rule, latency-tail upper bounds
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round(d.get('step_roofline', {}).get('frac', 0), 4), 'e2e_roi', d.get('e2e_roi', {}).get('value'))
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
timeout 600 python scripts/dw_bench.py 256 > gpurun_out/dw_bench.txt 2>&1; cat gpurun_out/dw_bench.txt | cut -c1-200
timeout 600 python scripts/gemm_bench2.py > gpurun_out/gemm_bench2.txt 2>&1; cat gpurun_out/gemm_bench2.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err | cut -c1-300; bench_line gpurun_out/bench.json
for m in 1 2 4 7; do
  TD3D_DBG_SKIP=$m timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_skip$m.json 2> gpurun_out/bench_skip$m.err; echo "skip$m rc=$?"; bench_line gpurun_out/bench_skip$m.json
done
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; bench_line gpurun_out/bench_infer.json
