set -x
mkdir -p gpurun_out
: > gpurun_out/r02_pytest_gpu_4.log
for f in tests/test_gpu_field.py tests/test_gpu_wide_heads.py tests/test_gpu_train.py tests/test_gpu_baseline_sizes.py; do
  echo "=== $f" >> gpurun_out/r02_pytest_gpu_4.log
  timeout 400 python -m pytest $f -q -m gpu --timeout=150 --timeout-method=thread >> gpurun_out/r02_pytest_gpu_4.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_pytest_gpu_4.log
done
grep -E "^===|passed|failed|rc=|^FAILED|^ERROR|Timeout|^E  " gpurun_out/r02_pytest_gpu_4.log | cut -c1-200
for cfg in base_light large small; do
(timeout 300 python bench.py --config $cfg --no-cpu-baseline --no-extras > gpurun_out/r02_bench_${cfg}_d.json 2> gpurun_out/r02_bench_${cfg}_d.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_${cfg}_d.err)
python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_${cfg}_d.json"))
print("$cfg", d["ms_per_step"], d["value"])
for k,v in d["extras"]["kernels"].items(): print("   ", k, v["ms_per_step"])
PY
done
