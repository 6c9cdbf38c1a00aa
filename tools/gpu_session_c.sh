set -x
RPK_TUNE=k2=beats4 timeout 200 python tools/k2_tune.py --iters 8 --slots 16777216 --strides 16,32 --out gpurun_out/r2_k2_tune_beats4.json 2>&1 | tail -4
timeout 200 python tools/k2_tune.py --iters 8 --slots 1000000,16777216 --strides 16,32 --out gpurun_out/r2_k2_tune_final.json 2>&1 | tail -8
for tool in memcheck racecheck synccheck; do
  ( time timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py ) > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2_sanitizer_$tool.log
  tail -6 gpurun_out/r2_sanitizer_$tool.log
done
