#!/bin/bash
# One-box evidence run (B200 x1): bench lines of both arms, the launch list of the bench, ncu --set full captures of the kernels
# that changed this round, the oracle-at-scale listings.  `bash scripts/run_evidence.sh <tag>` -> gpurun_out/<tag>_*.
tag=${1:-r3}
out=gpurun_out
mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-frames 4096 > $out/${tag}_ncu_launches.log 2>&1
for mode in exit fixed; do
  flag=""; [ $mode = fixed ] && flag="--fixed"
  ncu --set full --clock-control none --import-source on -k regex:vamp_quad -s 2 -c 1 -f -o $out/${tag}_vamp_quad_$mode \
      python scripts/profile_c3.py $flag > $out/${tag}_ncu_quad_$mode.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:vamp_fast -s 2 -c 1 -f -o $out/${tag}_vamp_fast_$mode \
      python scripts/profile_vamp.py $flag > $out/${tag}_ncu_vfast_$mode.log 2>&1
done
python -m pytest tests/test_gpu_oracle_scale.py -q -s -m gpu > $out/${tag}_oracle_scale.log 2>&1
tail -3 $out/${tag}_oracle_scale.log
