set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/b8.err | grep "^{" > gpurun_out/r2_bench_n8.json; tail -2 gpurun_out/b8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --stream-seconds 0 2>gpurun_out/b4.err | grep "^{" > gpurun_out/r2_bench_n4.json; tail -2 gpurun_out/b4.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/gather_probe_multi.py --iters 20 > gpurun_out/r2_gather_probe_n8.json 2> gpurun_out/probe_err.log; grep "rank 0" gpurun_out/r2_gather_probe_n8.json
timeout 300 python -m pytest tests/test_multi_gpu.py -q --timeout 150 --timeout-method thread > gpurun_out/r2_pytest_multi_gpu_n8.log 2>&1; tail -2 gpurun_out/r2_pytest_multi_gpu_n8.log
