set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 150 --timeout-method thread > gpurun_out/r2_pytest_gpu_n1.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_n1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_n1.err | grep "^{" > gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/r2_bench_reference_arm.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-k2-sweep --launch eager --stream-seconds 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_select_persist --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_k1_persist96_1m python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-k2-sweep --launch eager --stream-seconds 0 --no-parity > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_status_stream --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2_k2_stream32_16m python tools/k2_tune.py --iters 3 --slots 16777216 --strides 32 --out gpurun_out/k2_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pod|k_select" -c 24 --csv --log-file gpurun_out/r2_launches_select_125k.csv python tools/k1_tune.py --iters 3 --pods 125000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pod|k_select" -c 24 --csv --log-file gpurun_out/r2_launches_select_1m.csv python tools/k1_tune.py --iters 3 --pods 1000000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pod_scatter --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_k0_scatter_1m python tools/k1_tune.py --iters 3 --pods 1000000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pod_classify --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_k0_classify_1m python tools/k1_tune.py --iters 3 --pods 1000000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_select_persist --launch-skip 3 --launch-count 1 -f -o gpurun_out/r2_k1_persist96_125k python tools/k1_tune.py --iters 3 --pods 125000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
ls -la gpurun_out | tail -12
