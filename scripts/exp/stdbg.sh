for d in 0 32; do echo "dbg=$d"; AMPSM_ST_DEBUG=$d timeout 300 python scripts/bench_scamp.py --frames 1024 --fixed --reps 2 2>&1 | tail -1; done
