mkdir -p gpurun_out
: > gpurun_out/r02_pytest_gpu_5.log
for f in tests/test_gpu_raymarch.py tests/test_gpu_train.py tests/test_gpu_x_infer_loop.py tests/test_gpu_baseline_sizes.py; do
  echo "=== $f" >> gpurun_out/r02_pytest_gpu_5.log
  timeout 400 python -m pytest $f -q -m gpu --timeout=150 --timeout-method=thread >> gpurun_out/r02_pytest_gpu_5.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_pytest_gpu_5.log
done
grep -E "^===|passed|failed|rc=|^FAILED|^ERROR|Timeout|^E  " gpurun_out/r02_pytest_gpu_5.log | cut -c1-220
(timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02_bench_base_light_e.json 2> gpurun_out/r02_bench_base_light_e.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_base_light_e.err)
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_base_light_e.json"))
print("base_light", d["ms_per_step"], d["value"], "launches", d["gpu_launches"])
for k,v in d["extras"]["kernels"].items(): print("   ", k, v["ms_per_step"])
PY
(timeout 300 python bench.py --mode render --steps 3 --warmup 1 > gpurun_out/r02_bench_render_b.json 2> gpurun_out/r02_bench_render_b.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_render_b.err; head -c 600 gpurun_out/r02_bench_render_b.json)
