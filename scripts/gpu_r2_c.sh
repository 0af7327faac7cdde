# round 2, call C (2 GPUs): EfficientNet-B3 after the pack-table fix, inference bench error text, depthwise micro-benchmark,
# data-parallel test + N=2 bench (graph teardown fix), EfficientNet-B0 bench
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run_t() { name=$1; shift; timeout 1200 python -m pytest "$@" -q 2>&1 | tail -40 > gpurun_out/t_$name.log; echo "== $name"; tail -n 12 gpurun_out/t_$name.log | cut -c1-500; }
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round(d['step_roofline']['frac'], 4))
    kk = d.get('kernel_kinds') or {}
    if kk:
        key = 'ms_per_step' if 'ms_per_step' in next(iter(kk.values())) else 'ms_per_chunk'
        for k, v in sorted(kk.items(), key=lambda kv: -kv[1][key])[:12]:
            print(f"  {k:16s} {v[key]:8.3f} ms  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
run_t effnet tests/test_gpu_effnet.py
run_t dwx tests/test_gpu_kernels.py -k "explicit"
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; grep -v "^frame\|^$" gpurun_out/bench_infer.err | head -30 | cut -c1-400; bench_line gpurun_out/bench_infer.json
timeout 600 python bench.py --mode infer --steps 5 --warmup 3 --infer-batch 1024 --skip-cpu > gpurun_out/bench_infer1k.json 2> gpurun_out/bench_infer1k.err; echo "infer1k rc=$?"; grep -v "^frame\|^$" gpurun_out/bench_infer1k.err | head -12 | cut -c1-400; bench_line gpurun_out/bench_infer1k.json
timeout 600 python scripts/dw_bench.py 256 > gpurun_out/dw_bench.txt 2>&1; cat gpurun_out/dw_bench.txt | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -30 > gpurun_out/t_dp.log; tail -n 25 gpurun_out/t_dp.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench2_graph.json 2> gpurun_out/bench2_graph.err; echo "bench2 graph rc=$?"; bench_line gpurun_out/bench2_graph.json
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; tail -n 3 gpurun_out/bench_b0.err | cut -c1-300; bench_line gpurun_out/bench_b0.json
timeout 900 python bench.py --workload effnet_b3 --steps 5 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_b3.json 2> gpurun_out/bench_b3.err; echo "b3 rc=$?"; tail -n 3 gpurun_out/bench_b3.err | cut -c1-300; bench_line gpurun_out/bench_b3.json
timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench1 rc=$?"; bench_line gpurun_out/bench1.json
ls -la gpurun_out/ | tail -12
