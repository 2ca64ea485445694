mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_train.py -q -m gpu --timeout=100 --timeout-method=thread 2>&1 | tail -3
(timeout 100 python bench.py --no-cpu-baseline --no-extras --steps 10 --warmup 3 > gpurun_out/r02_bench_base_light_last.json 2> gpurun_out/r02_bench_base_light_last.err; echo rc=$?)
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_base_light_last.json")); print(d["ms_per_step"], d["value"])
PY
