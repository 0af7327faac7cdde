# round 2, call D (2 GPUs): after the tcgen05 accumulator hand-off fix
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run_t() { name=$1; shift; timeout 1200 python -m pytest "$@" -q 2>&1 | tail -60 > gpurun_out/t_$name.log; echo "== $name"; tail -n 40 gpurun_out/t_$name.log | cut -c1-400; }
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
run_t dp tests/test_gpu_dp.py
run_t tc tests/test_gpu_tc.py tests/test_gpu_roi.py
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; grep -v "^frame\|^$" gpurun_out/bench_infer.err | head -12 | cut -c1-300; bench_line gpurun_out/bench_infer.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench2_graph.json 2> gpurun_out/bench2_graph.err; echo "bench2 graph rc=$?"; grep -v "^frame\|^$" gpurun_out/bench2_graph.err | head -12 | cut -c1-300; bench_line gpurun_out/bench2_graph.json
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; grep -v "^frame\|^$" gpurun_out/bench_b0.err | head -12 | cut -c1-300; bench_line gpurun_out/bench_b0.json
timeout 900 python bench.py --workload effnet_b3 --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_b3.json 2> gpurun_out/bench_b3.err; echo "b3 rc=$?"; grep -v "^frame\|^$" gpurun_out/bench_b3.err | head -12 | cut -c1-300; bench_line gpurun_out/bench_b3.json
timeout 600 python scripts/dw_bench.py 256 > gpurun_out/dw_bench.txt 2>&1; cat gpurun_out/dw_bench.txt | cut -c1-200
TD3D_DWC_TALL=0 timeout 600 python scripts/dw_bench.py 256 0,1,2,3,4,5 > gpurun_out/dw_bench_r2.txt 2>&1; cat gpurun_out/dw_bench_r2.txt | cut -c1-200
TD3D_DW_BWD_FUSED=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "fused rc=$?"; bench_line gpurun_out/bench_fused.json
timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench1 rc=$?"; bench_line gpurun_out/bench1.json
timeout 300 ncu --set full --import-source off --clock-control none -k regex:"dwc_bwd|dwc_fwd" -c 6 -o gpurun_out/dwc_small python scripts/dw_bench.py 256 1,2,13 > gpurun_out/ncu_dwc.out 2>&1; python scripts/ncu_full_table.py gpurun_out/dwc_small.ncu-rep > gpurun_out/dwc_ncu_table.txt 2>&1; cat gpurun_out/dwc_ncu_table.txt | cut -c1-260; rm -f gpurun_out/dwc_small.ncu-rep
ls -la gpurun_out/ | tail -8
