mkdir -p gpurun_out
N=4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 200 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/r02_bench_n4_weak_prio.json 2> gpurun_out/r02_bench_n4_weak_prio.err; echo "rc=$?")
python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02_bench_n4_weak_prio.json").read().splitlines() if l.startswith("{")][-1]
open("gpurun_out/r02_bench_n4_weak_prio.json","w").write(line+"\n")
d=json.loads(line); print("weak", d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
(timeout 200 $TR --master-port 29501 profiles/check_peer_exchange.py base_light > gpurun_out/r02_peer_check_n4.log 2>&1; echo "rc=$?"; grep -E "peer_exchange_check|Error|error" gpurun_out/r02_peer_check_n4.log | head -5)
