set -x
timeout 900 python -m pytest tests/test_status_gpu.py tests/test_select_gpu.py -m gpu -x -q --timeout 150 --timeout-method thread 2>&1 | tail -4
timeout 200 python tools/k2_tune.py --iters 8 --slots 16777216 --strides 16,32 --out gpurun_out/r2_k2_tune_beats.json 2>&1 | tail -4
RPK_TUNE=k2=units timeout 200 python tools/k2_tune.py --iters 8 --slots 16777216 --strides 16,32 --out gpurun_out/r2_k2_tune_units.json 2>&1 | tail -4
timeout 300 python tools/k1_tune.py --iters 20 --pods 125000,1000000 --variants "|k0ctas=100000" --out gpurun_out/r2_k1_tune_k0b.json 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pod|k_select" -c 24 --csv --log-file gpurun_out/r2_launches_select_1m.csv python tools/k1_tune.py --iters 3 --pods 1000000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_pod|k_select" -c 24 --csv --log-file gpurun_out/r2_launches_select_125k.csv python tools/k1_tune.py --iters 3 --pods 125000 --variants "" --out gpurun_out/k1_ncu_dummy.json > /dev/null 2>&1
