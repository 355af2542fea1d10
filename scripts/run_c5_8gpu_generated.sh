#!/bin/bash
# BASELINE config 5 with the frames drawn inside the SVD kernel (simulate.py --generator kernel: the channel matrices never
# exist in HBM): VAMP on Kronecker-correlated 64 x 32 channels, 1e8 frames per Eb/N0 point sharded over the ranks of one box.
# Usage: gpurun --gpus 8 -- bash scripts/run_c5_8gpu_generated.sh [frames]
N=${NGPU:-8}
FR=${1:-100000000}
out=gpurun_out/c5gen_${N}gpu
mkdir -p $out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29521 simulate.py --generator kernel --alg vamp --channel kronecker --rho-t 0.7 --rho-r 0.7 --alphabet QPSK --frames $FR --ebn0-start 4 --ebn0-final 12 --ebn0-step 4 --path $out/rho07 > $out/rho07.log 2>&1
tail -4 $out/rho07.log
run 29522 simulate.py --generator kernel --alg vamp --channel kronecker --rho-t 0.9 --rho-r 0.9 --alphabet QPSK --frames $FR --ebn0-start 10 --ebn0-final 18 --ebn0-step 8 --path $out/rho09 > $out/rho09.log 2>&1
tail -3 $out/rho09.log
