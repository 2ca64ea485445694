mkdir -p gpurun_out
N=${NGPU:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
parse() { python - "$1" "$2" <<'PY'
import json,sys
try:
    line=[l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1]
    d=json.loads(line); print(sys.argv[2], d["n_gpus"], round(d["ms_per_step"],4), round(d["value"]), d.get("scaling"), d["config"].get("parallelism","")[-90:], "|", d["config"].get("launch",""), "| e2e", round(d["e2e"]["value"]))
except Exception as e: print(sys.argv[2], "no json", e)
PY
}
(timeout 300 $TR --master-port 29506 profiles/check_render_sharded.py 4096 > gpurun_out/r02_render_check_n${N}_v3.log 2>&1; echo "rc=$?"; grep -E "render_sharded_check|Error|error" gpurun_out/r02_render_check_n${N}_v3.log | head -5)
(timeout 300 $TR --master-port 29507 bench.py --gpus $N --mode render --steps 5 --warmup 2 > gpurun_out/r02_bench_render_n${N}_v3.json 2> gpurun_out/r02_bench_render_n${N}_v3.err; echo "rc=$?"); parse gpurun_out/r02_bench_render_n${N}_v3.json render
(timeout 240 $TR --master-port 29502 bench.py --gpus $N --no-cpu-baseline --no-extras > gpurun_out/r02_bench_n${N}_weak_v3.json 2> gpurun_out/r02_bench_n${N}_weak_v3.err; echo "rc=$?"); parse gpurun_out/r02_bench_n${N}_weak_v3.json weak_final
(timeout 240 $TR --master-port 29505 bench.py --gpus $N --no-cpu-baseline --no-extras --scaling strong > gpurun_out/r02_bench_n${N}_strong_v3.json 2> gpurun_out/r02_bench_n${N}_strong_v3.err; echo "rc=$?"); parse gpurun_out/r02_bench_n${N}_strong_v3.json strong_final
