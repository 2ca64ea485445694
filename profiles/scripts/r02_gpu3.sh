set -x
mkdir -p gpurun_out
: > gpurun_out/r02_pytest_gpu_3.log
for f in tests/test_gpu_wide_heads.py tests/test_gpu_umma.py tests/test_gpu_baseline_sizes.py tests/test_gpu_train.py tests/test_gpu_encoder.py tests/test_gpu_field.py tests/test_gpu_raymarch.py tests/test_gpu_x_feeder.py tests/test_gpu_x_infer_loop.py; do
  echo "=== $f" >> gpurun_out/r02_pytest_gpu_3.log
  timeout 500 python -m pytest $f -q -m gpu --timeout=200 --timeout-method=thread >> gpurun_out/r02_pytest_gpu_3.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_pytest_gpu_3.log
done
grep -E "^===|passed|failed|rc=|^FAILED|^ERROR|Timeout" gpurun_out/r02_pytest_gpu_3.log
(timeout 300 python bench.py --config large --no-cpu-baseline --no-extras > gpurun_out/r02_bench_large_c.json 2> gpurun_out/r02_bench_large_c.err; echo rc=$?; tail -c 500 gpurun_out/r02_bench_large_c.err)
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_large_c.json"))
print(d["ms_per_step"], d["value"])
for k,v in d["extras"]["kernels"].items(): print(k, v["ms_per_step"])
PY
