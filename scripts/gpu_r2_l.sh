# round 2, call L (1 GPU): programmatic dependent launch on every kernel of the chain
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
bench_line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'step_frac', round(d.get('step_roofline', {}).get('frac', 0), 4), 'e2e_roi', d.get('e2e_roi', {}).get('value'))
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
( time timeout 2400 python -m pytest tests -q -m gpu -x ) > gpurun_out/t_all.log 2>&1; tail -n 8 gpurun_out/t_all.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "pdl rc=$?"; tail -n 3 gpurun_out/bench_pdl.err | cut -c1-300; bench_line gpurun_out/bench_pdl.json
TD3D_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "nopdl rc=$?"; bench_line gpurun_out/bench_nopdl.json
timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile --no-graph > gpurun_out/bench_pdl_eager.json 2> gpurun_out/bench_pdl_eager.err; echo "pdl eager rc=$?"; bench_line gpurun_out/bench_pdl_eager.json
TD3D_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-infer --skip-cpu --skip-profile --no-graph > gpurun_out/bench_nopdl_eager.json 2> gpurun_out/bench_nopdl_eager.err; echo "nopdl eager rc=$?"; bench_line gpurun_out/bench_nopdl_eager.json
timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; echo "infer rc=$?"; tail -n 3 gpurun_out/bench_infer.err | cut -c1-300; bench_line gpurun_out/bench_infer.json
TD3D_PDL=0 timeout 600 python bench.py --mode infer --steps 10 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_infer_nopdl.json 2> gpurun_out/bench_infer_nopdl.err; echo "infer nopdl rc=$?"; bench_line gpurun_out/bench_infer_nopdl.json
timeout 900 python bench.py --workload effnet_b0 --steps 10 --warmup 3 --skip-cpu --skip-profile > gpurun_out/bench_b0.json 2> gpurun_out/bench_b0.err; echo "b0 rc=$?"; bench_line gpurun_out/bench_b0.json
