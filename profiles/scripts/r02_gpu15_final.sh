set -x
mkdir -p gpurun_out
: > gpurun_out/r02_pytest_gpu_final.log
for f in tests/test_gpu_*.py; do
  echo "=== $f" >> gpurun_out/r02_pytest_gpu_final.log
  timeout 400 python -m pytest $f -q -m gpu --timeout=200 --timeout-method=thread >> gpurun_out/r02_pytest_gpu_final.log 2>&1
  echo "rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
done
grep -E "^===|passed|failed|rc=|^FAILED|^ERROR|Timeout|^E  " gpurun_out/r02_pytest_gpu_final.log | cut -c1-200
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke.log)
(timeout 500 python bench.py > gpurun_out/r02_bench_base_light_final.json 2> gpurun_out/r02_bench_base_light_final.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_base_light_final.err)
(timeout 300 python bench.py --config small --no-cpu-baseline > gpurun_out/r02_bench_small_final.json 2> gpurun_out/r02_bench_small_final.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_small_final.err)
(timeout 300 python bench.py --config large --no-cpu-baseline > gpurun_out/r02_bench_large_final.json 2> gpurun_out/r02_bench_large_final.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_large_final.err)
(timeout 300 python bench.py --mode render --steps 3 --warmup 1 > gpurun_out/r02_bench_render_n1_final.json 2> gpurun_out/r02_bench_render_n1_final.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_render_n1_final.err)
(timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo rc=$?; tail -c 300 gpurun_out/r02_bench_reference_arm.err; head -c 700 gpurun_out/r02_bench_reference_arm.json)
(TNL_PREFETCH=0 TNL_STEPS=2 timeout 500 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r02_step_full_final -f python profiles/prof_step.py > gpurun_out/r02_ncu_step_final.log 2>&1; echo rc=$?; tail -2 gpurun_out/r02_ncu_step_final.log)
ncu -i gpurun_out/r02_step_full_final.ncu-rep --page raw --csv > gpurun_out/r02_step_raw_final.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_step_full_final.ncu-rep > gpurun_out/r02_ncu_step_final.txt 2>&1
rm -f gpurun_out/r02_step_full_final.ncu-rep
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_launches_bench.log 2>&1; echo rc=$?)
(TNL_CONFIG=large TNL_PREFETCH=0 TNL_STEPS=2 timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:k_mlp_tc -o gpurun_out/r02_mlp128 -f python profiles/prof_step.py > gpurun_out/r02_ncu_mlp128.log 2>&1; echo rc=$?)
python profiles/ncu_summary.py gpurun_out/r02_mlp128.ncu-rep > gpurun_out/r02_ncu_mlp128.txt 2>&1
rm -f gpurun_out/r02_mlp128.ncu-rep
python - <<'PY'
import json
for n in ["base_light","small","large"]:
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{n}_final.json"))
        print(n, round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
    except Exception as e: print(n, "ERR", e)
try:
    d=json.load(open("gpurun_out/r02_bench_render_n1_final.json")); print("render", d["ms_per_step"], d["value"])
except Exception as e: print("render ERR", e)
PY
