mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m | head -12
(timeout 400 $TR --master-port 29501 profiles/check_peer_exchange.py base_light > gpurun_out/r02_peer_check_n$N.log 2>&1; echo "rc=$?"; grep -E "peer_exchange_check|Error|error|Traceback" gpurun_out/r02_peer_check_n$N.log | head -20; tail -5 gpurun_out/r02_peer_check_n$N.log)
for ex in auto nccl; do
(timeout 300 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras --exchange $ex > gpurun_out/r02_bench_n${N}_weak_$ex.json 2> gpurun_out/r02_bench_n${N}_weak_$ex.err; echo "rc=$?"; tail -c 600 gpurun_out/r02_bench_n${N}_weak_$ex.err | tail -5)
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}_weak_$ex.json")); print("$ex weak", d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["parallelism"][-160:], d["config"]["launch"])
except Exception as e: print("no json", e)
PY
done
(timeout 300 $TR --master-port 29503 bench.py --gpus $N --no-cpu-baseline --no-extras --scaling strong > gpurun_out/r02_bench_n${N}_strong.json 2> gpurun_out/r02_bench_n${N}_strong.err; echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_n${N}_strong.err | tail -3)
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_n${N}_strong.json")); print("strong", d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["rays_per_gpu"])
    for k,v in d["extras"]["kernels"].items(): print("   ", k, v["ms_per_step"])
except Exception as e: print("no json", e)
PY
