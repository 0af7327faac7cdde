set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"apply_xform" --launch-skip 66 -c 3 -o gpurun_out/ew_ax python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_ax.out 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"act_bwd_stats|affine2" --launch-skip 140 -c 10 -o gpurun_out/ew_ab python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-profile > gpurun_out/ncu_ab.out 2>&1
ls -la gpurun_out | grep ew_
