mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 240 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/r02_bench_n8_weak_final.json 2> gpurun_out/r02_bench_n8_weak_final.err; echo "rc=$?")
python - <<'PY'
import json
line=[l for l in open("gpurun_out/r02_bench_n8_weak_final.json").read().splitlines() if l.startswith("{")][-1]
open("gpurun_out/r02_bench_n8_weak_final.json","w").write(line+"\n")
d=json.loads(line); print("weak", d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]))
PY
