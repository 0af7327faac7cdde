set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"ww_" --launch-skip 60 -c 14 -o gpurun_out/ww_full python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_ww.out 2>&1
ls -la gpurun_out/
