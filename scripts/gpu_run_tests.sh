set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu 2>&1 | tail -300 > gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu 2>&1 | tail -300 > gpurun_out/t_tc.log
timeout 1200 python -m pytest tests/test_gpu_model.py -q -m gpu 2>&1 | tail -600 > gpurun_out/t_model.log
for f in gpurun_out/t_kernels.log gpurun_out/t_tc.log gpurun_out/t_model.log; do tail -n 3 $f; done
