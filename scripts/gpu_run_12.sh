set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
grep -E "passed|failed|Error" gpurun_out/t_all.log | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-profile > gpurun_out/bench_overlap.json 2> gpurun_out/bench_overlap.err
tail -n 3 gpurun_out/bench_overlap.err; cut -c1-220 gpurun_out/bench_overlap.json
TD3D_OVERLAP=0 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-profile > gpurun_out/bench_nooverlap.json 2> gpurun_out/bench_nooverlap.err
cut -c1-220 gpurun_out/bench_nooverlap.json
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-profile --no-graph > gpurun_out/bench_overlap_eager.json 2> gpurun_out/bench_overlap_eager.err
cut -c1-220 gpurun_out/bench_overlap_eager.json
