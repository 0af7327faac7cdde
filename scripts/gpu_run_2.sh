# dw2 correctness + short bench + ncu full captures (run under gpurun)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "depthwise" 2>&1 | tail -15 > gpurun_out/t_dw.log
tail -5 gpurun_out/t_dw.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/t_all.log
grep -E "passed|failed" gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -n 5 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s 1 -c 3 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_gemm.out 2>&1
tail -n 3 gpurun_out/ncu_gemm.out | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:d2_ -c 8 -o gpurun_out/prof_dw python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_dw.out 2>&1
tail -n 3 gpurun_out/ncu_dw.out | cut -c1-300
ls -la gpurun_out
