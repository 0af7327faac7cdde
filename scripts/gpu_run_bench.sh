# tests + bench + launch list (run under gpurun)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -300 > gpurun_out/t_all.log
grep -E "passed|failed" gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -n 15 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_bench.out 2>&1
tail -n 3 gpurun_out/ncu_bench.out | cut -c1-400
wc -l gpurun_out/launches.csv
