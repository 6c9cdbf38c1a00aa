set -x
for tool in memcheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py ) > gpurun_out/r2_sanitizer2_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_sanitizer2_$tool.log
  tail -7 gpurun_out/r2_sanitizer2_$tool.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | grep "^{" > gpurun_out/r2_bench_n2.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/gather_probe_multi.py --iters 20 > gpurun_out/r2_gather_probe_n2.json 2> gpurun_out/probe_err.log; grep "rank 0" gpurun_out/r2_gather_probe_n2.json
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q --timeout 120 --timeout-method thread 2>&1 | tail -3
