# round 2, call X (1 GPU): pipeline timeline of the NT GEMM (all-run-time instance)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python scripts/gemm_bench.py > gpurun_out/gemm_bench_timeline.txt 2>&1; tail -n 60 gpurun_out/gemm_bench_timeline.txt | cut -c1-250
