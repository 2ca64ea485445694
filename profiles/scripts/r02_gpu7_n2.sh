mkdir -p gpurun_out
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
(timeout 400 $TR --master-port 29501 profiles/check_peer_exchange.py base_light > gpurun_out/r02_peer_check_n$N.log 2>&1; echo "rc=$?"; grep -E "peer_exchange_check|Error|error" gpurun_out/r02_peer_check_n$N.log | head -20)
