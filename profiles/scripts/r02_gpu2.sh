set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2.log; tail -40 gpurun_out/r02_pytest_gpu_2.log)
(timeout 400 python bench.py --config large --no-cpu-baseline > gpurun_out/r02_bench_large_b.json 2> gpurun_out/r02_bench_large_b.err; echo rc=$?; tail -c 800 gpurun_out/r02_bench_large_b.err; head -c 2500 gpurun_out/r02_bench_large_b.json)
