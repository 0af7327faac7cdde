# round 2, call B (1 GPU): parity of the new kernels + A/B benches
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run_t() { name=$1; shift; timeout 1200 python -m pytest "$@" -q 2>&1 | tail -40 > gpurun_out/t_$name.log; echo "== $name"; tail -n 12 gpurun_out/t_$name.log | cut -c1-500; }
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round(d['step_roofline']['frac'], 4))
    kk = d['kernel_kinds']; key = 'ms_per_step' if 'ms_per_step' in next(iter(kk.values())) else 'ms_per_chunk'
    for k, v in sorted(kk.items(), key=lambda kv: -kv[1][key])[:12]:
        print(f"  {k:16s} {v[key]:8.3f} ms  {v['gbs']:8.1f} GB/s")
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
run_t dw tests/test_gpu_kernels.py -x -k "depthwise or stem"
run_t model tests/test_gpu_model.py -x
TD3D_DWC_PF=6 timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --dump-launches gpurun_out/launches_pf6.csv > gpurun_out/bench_pf6.json 2> gpurun_out/bench_pf6.err; echo "rc=$?"; bench_line gpurun_out/bench_pf6.json
grep "dw_bwd\|dw_fwd" gpurun_out/launches_pf6.csv | head -40
run_t infer tests/test_gpu_infer.py
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; tail -n 3 gpurun_out/bench_infer.err | cut -c1-400; bench_line gpurun_out/bench_infer.json
run_t effnet tests/test_gpu_effnet.py
run_t surface tests/test_gpu_surface.py
TD3D_DW_FWD_CW=1 timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --dump-launches gpurun_out/launches_fwdcw.csv > gpurun_out/bench_fwdcw.json 2> gpurun_out/bench_fwdcw.err; echo "rc=$?"; bench_line gpurun_out/bench_fwdcw.json
grep "dw_fwd" gpurun_out/launches_fwdcw.csv | head -20
for pf in 0 12; do
  TD3D_DWC_PF=$pf timeout 300 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf$pf.err; echo "pf=$pf rc=$?"; bench_line gpurun_out/bench_pf$pf.json
done
run_t rest tests/test_gpu_kernels.py tests/test_gpu_tc.py -k "not depthwise and not stem"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"dwc_bwd|dwc_fwd" --launch-skip 30 -c 24 -o gpurun_out/dwc_full python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile --skip-infer > gpurun_out/ncu_dwc.out 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | tail -12
