# round 2, call A (2 GPUs): data-parallel bring-up + big-M GEMM reproducer
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
for args in "1605632 64 16 128" "12845056 64 16 0" "12845056 64 16 1024" "12845056 16 16 1024" "6422528 64 16 512"; do
  timeout 300 python tests/tc_bigm.py $args > gpurun_out/bigm_$(echo $args | tr ' ' '_').log 2>&1; echo "rc=$? $args"; tail -n 2 gpurun_out/bigm_$(echo $args | tr ' ' '_').log | cut -c1-400
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 --log-dir gpurun_out/dp_logs --redirects 3 --tee 3 tests/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"
tail -n 30 gpurun_out/dp_check.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 --log-dir gpurun_out/b2_logs --redirects 3 --tee 3 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph --skip-infer > gpurun_out/bench2_nograph.log 2>&1; echo "bench2 nograph rc=$?"
tail -n 5 gpurun_out/bench2_nograph.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 --log-dir gpurun_out/b2g_logs --redirects 3 --tee 3 bench.py --gpus 2 --steps 20 --warmup 5 --skip-infer > gpurun_out/bench2_graph.log 2>&1; echo "bench2 graph rc=$?"
tail -n 5 gpurun_out/bench2_graph.log | cut -c1-1500
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --skip-infer --skip-cpu > gpurun_out/bench1.log 2>&1; echo "bench1 rc=$?"
tail -n 2 gpurun_out/bench1.log | cut -c1-600
