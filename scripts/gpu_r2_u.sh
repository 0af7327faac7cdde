# round 2, call U (1 GPU): ncu --set full on the NT GEMM (wide N, with / without statistics)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc --launch-skip 2 -c 1 -f -o gpurun_out/gemm_240_40_st python scripts/one_gemm.py 200704 240 40 32 > gpurun_out/ncu_u1.log 2>&1; tail -n 2 gpurun_out/ncu_u1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc --launch-skip 2 -c 1 -f -o gpurun_out/gemm_240_40_ns python scripts/one_gemm.py 200704 240 40 0 > gpurun_out/ncu_u2.log 2>&1; tail -n 2 gpurun_out/ncu_u2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc --launch-skip 2 -c 1 -f -o gpurun_out/gemm_64_16_st python scripts/one_gemm.py 3211264 64 16 32 > gpurun_out/ncu_u3.log 2>&1; tail -n 2 gpurun_out/ncu_u3.log
ls -la gpurun_out/*.ncu-rep
