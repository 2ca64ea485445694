set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1.log; tail -15 gpurun_out/r02_pytest_gpu_1.log)
(timeout 600 python bench.py > gpurun_out/r02_bench_base_light_a.json 2> gpurun_out/r02_bench_base_light_a.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_base_light_a.err)
(timeout 300 python bench.py --config small --no-cpu-baseline > gpurun_out/r02_bench_small_a.json 2> gpurun_out/r02_bench_small_a.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_small_a.err)
(timeout 300 python bench.py --config large --no-cpu-baseline > gpurun_out/r02_bench_large_a.json 2> gpurun_out/r02_bench_large_a.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_large_a.err)
(timeout 400 python bench.py --mode render --steps 3 --warmup 1 > gpurun_out/r02_bench_render_a.json 2> gpurun_out/r02_bench_render_a.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_render_a.err)
(TNL_PREFETCH=0 TNL_STEPS=2 timeout 600 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r02_step_full -f python profiles/prof_step.py > gpurun_out/r02_ncu_step.log 2>&1; echo rc=$?; tail -3 gpurun_out/r02_ncu_step.log)
ncu -i gpurun_out/r02_step_full.ncu-rep --page raw --csv > gpurun_out/r02_step_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_step_full.ncu-rep > gpurun_out/r02_ncu_step_a.txt 2>&1
ls -la gpurun_out/ | head -30
rm -f gpurun_out/r02_step_full.ncu-rep
for f in gpurun_out/r02_bench_*_a.json; do echo $f; head -c 1500 $f; echo; done
