#!/bin/bash
# BASELINE config 5 on one 8-GPU box: VAMP on Kronecker-correlated channels, 1e8 frames per Eb/N0 point sharded over the ranks,
# one NCCL all-reduce of the counter block per point; then the headline bench at N = 8 with and without the NUMA binding of
# the host path.  Usage: gpurun --gpus 8 -- bash scripts/run_c5_8gpu.sh [frames]
N=${NGPU:-8}
FR=${1:-100000000}
out=gpurun_out/c5_${N}gpu
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
lscpu > $out/lscpu.txt 2>&1
(numactl -H || cat /sys/devices/system/node/node*/cpulist) > $out/numa.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 simulate.py --alg vamp --channel kronecker --rho-t 0.7 --rho-r 0.7 --alphabet QPSK --frames $FR --ebn0-start 4 --ebn0-final 12 --ebn0-step 4 --path $out/rho07 > $out/rho07.log 2>&1
tail -4 $out/rho07.log
run 29512 simulate.py --alg vamp --channel kronecker --rho-t 0.9 --rho-r 0.9 --alphabet QPSK --frames $FR --ebn0-start 10 --ebn0-final 18 --ebn0-step 8 --path $out/rho09 > $out/rho09.log 2>&1
tail -3 $out/rho09.log
short="--no-fixed-t --vamp-frames 0 --c3-frames 0 --scamp-frames 0 --c1-frames 0 --c128-frames 0 --no-cpu-baseline"
run 29513 bench.py --gpus $N --steps 5 --warmup 3 $short > $out/bench_numa.json 2> $out/bench_numa.err
AMPSM_NO_NUMA_BIND=1 run 29514 bench.py --gpus $N --steps 5 --warmup 3 $short > $out/bench_nonuma.json 2> $out/bench_nonuma.err
python - <<PY
import json
for f in ("bench_numa","bench_nonuma"):
    try:
        d=json.load(open("$out/%s.json"%f)); print(f, "value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]), d["e2e"].get("numa"))
    except Exception as e: print(f, "failed", e)
PY
