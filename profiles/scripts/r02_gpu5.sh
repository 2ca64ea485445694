mkdir -p gpurun_out
timeout 200 python profiles/bench_umma.py > gpurun_out/r02_bench_umma.log 2>&1; tail -60 gpurun_out/r02_bench_umma.log
timeout 200 python profiles/prof_mlp_tc.py > gpurun_out/r02_prof_mlp_tc.log 2>&1; tail -30 gpurun_out/r02_prof_mlp_tc.log
