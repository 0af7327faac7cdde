# Weak scaling of the data-parallel train step on one 8-GPU box:  gpurun --gpus 8 -- 'bash scripts/gpu_scale.sh'
# configs[1] (MobileNetV3-large, 256 crops / GPU) at N = 8, 4, 2, 1 and configs[2] (EfficientNet-B0, 512 / GPU) at N = 8, 1.
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -8
line() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(sys.argv[1], 'n', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], 'parse failed', e)
PY
}
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "N=$n rc=$?"; grep -v "^frame\|^$\|OMP_NUM\|^\*\*\*" gpurun_out/scale_n$n.err | head -8 | cut -c1-300; line gpurun_out/scale_n$n.json
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --skip-infer --skip-cpu --skip-profile > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "N=1 rc=$?"; line gpurun_out/scale_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29628 bench.py --gpus 8 --workload effnet_b0 --steps 10 --warmup 3 --skip-infer --skip-cpu --skip-profile > gpurun_out/scale_b0_n8.json 2> gpurun_out/scale_b0_n8.err; echo "b0 N=8 rc=$?"; grep -v "^frame\|^$\|OMP_NUM\|^\*\*\*" gpurun_out/scale_b0_n8.err | head -8 | cut -c1-300; line gpurun_out/scale_b0_n8.json
timeout 600 python bench.py --gpus 1 --workload effnet_b0 --steps 10 --warmup 3 --skip-infer --skip-cpu --skip-profile > gpurun_out/scale_b0_n1.json 2> gpurun_out/scale_b0_n1.err; echo "b0 N=1 rc=$?"; line gpurun_out/scale_b0_n1.json
timeout 900 python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -30 > gpurun_out/t_dp.log; tail -n 6 gpurun_out/t_dp.log | cut -c1-400
