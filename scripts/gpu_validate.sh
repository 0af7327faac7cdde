# Full validation on the GPU box (1 GPU, or 2 for the NCCL data-parallel test):  gpurun [--gpus 2] -- 'bash scripts/gpu_validate.sh'
# GPU test suite, smoke(), every bench line (train / reference arm / inference / EfficientNet-B0 / B3 / 2-GPU), the per-layer micro-benchmarks
# and the ncu launch list of one step.  Everything lands in gpurun_out/; profiles/ holds the copies that are kept.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round((d.get('step_roofline') or {}).get('frac', 0), 4), 'e2e_roi', (d.get('e2e_roi') or {}).get('value'))
    kk = d.get('kernel_kinds') or {}
    if kk:
        key = 'ms_per_step' if 'ms_per_step' in next(iter(kk.values())) else 'ms_per_chunk'
        for k, v in sorted(kk.items(), key=lambda kv: -kv[1][key])[:14]:
            print(f"  {k:16s} {v[key]:8.3f} ms  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
( time timeout 2400 python -m pytest tests -q -m gpu ) > gpurun_out/t_all.log 2>&1; tail -n 14 gpurun_out/t_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --dump-launches gpurun_out/launches_eager.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.err | cut -c1-300; bench_line gpurun_out/bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -n 1 gpurun_out/bench_ref.json | cut -c1-400
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; bench_line gpurun_out/bench_infer.json
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; bench_line gpurun_out/bench_b0.json
timeout 900 python bench.py --workload effnet_b3 --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_b3.json 2> gpurun_out/bench_b3.err; echo "b3 rc=$?"; bench_line gpurun_out/bench_b3.json
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; bench_line gpurun_out/bench2.json
fi
timeout 600 python scripts/dw_bench.py 256 > gpurun_out/dw_bench.txt 2>&1; tail -n 17 gpurun_out/dw_bench.txt | cut -c1-200
timeout 600 python scripts/gemm_bench3.py > gpurun_out/gemm_bench3.txt 2>&1; tail -n 5 gpurun_out/gemm_bench3.txt | cut -c1-200
timeout 600 python scripts/gemm_bench_tn.py > gpurun_out/gemm_bench_tn.txt 2>&1; tail -n 3 gpurun_out/gemm_bench_tn.txt | cut -c1-200
timeout 600 python scripts/iou_bench.py > gpurun_out/iou_bench.txt 2>&1; cat gpurun_out/iou_bench.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 800 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile --skip-infer > gpurun_out/ncu_bench.out 2>&1
python scripts/ncu_step_summary.py gpurun_out/launches.csv gpurun_out/launches_summary.txt gpurun_out/ncu_traffic.json | tail -24
ls -la gpurun_out/ | tail -5
