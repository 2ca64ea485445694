mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29501 profiles/check_peer_exchange.py base_light > gpurun_out/r02_peer_check_n$N.log 2>&1; echo "rc=$?"; grep -E "peer_exchange_check|Error|error" gpurun_out/r02_peer_check_n$N.log | head -20)
for ex in auto; do
(timeout 300 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras --exchange $ex > gpurun_out/r02_bench_n${N}_weak_$ex.json 2> gpurun_out/r02_bench_n${N}_weak_$ex.err; echo "rc=$?"; tail -c 300 gpurun_out/r02_bench_n${N}_weak_$ex.err | tail -3)
python - <<PY
import json
try:
    line=[l for l in open("gpurun_out/r02_bench_n${N}_weak_$ex.json").read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line); print("$ex weak", d["n_gpus"], d["ms_per_step"], d["value"], d["config"]["parallelism"][-100:], d["config"]["launch"])
except Exception as e: print("no json", e)
PY
done
