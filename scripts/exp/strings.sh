run() { timeout 300 python scripts/bench_scamp.py --frames 1024 --fixed --reps 2 2>&1 | tail -1 | sed 's/.*ms\/call=\([0-9.]*\).*/ms\/call \1/'; }
echo "default"; run
for r in "2,2" "3,2" "2,3" "4,1"; do echo "mode1 rings $r"; AMPSM_ST_RINGS1=$r run; done
for r in "3,4" "4,3" "5,3" "6,2" "4,4" "2,2"; do echo "mode0 rings $r"; AMPSM_ST_RINGS0=$r run; done
