mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
parse() { python - "$1" "$2" <<'PY'
import json,sys
try:
    line=[l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
    open(sys.argv[1],"w").write(line+"\n")
    d=json.loads(line); print(sys.argv[2], d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), d.get("scaling"), d["config"].get("parallelism","")[-80:], "| e2e", round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[2], "no json", e)
PY
}
(timeout 240 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/r02_bench_n${N}_weak_final.json 2> gpurun_out/r02_bench_n${N}_weak_final.err; echo "rc=$?"); parse gpurun_out/r02_bench_n${N}_weak_final.json weak
(timeout 240 $TR --master-port 29505 bench.py --gpus $N --no-cpu-baseline --no-extras --scaling strong > gpurun_out/r02_bench_n${N}_strong_final.json 2> gpurun_out/r02_bench_n${N}_strong_final.err; echo "rc=$?"); parse gpurun_out/r02_bench_n${N}_strong_final.json strong
