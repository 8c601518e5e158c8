P='import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["value"],3), round(d["ms_per_step"],3))'
run() { echo -n "$* : "; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"; }
run WN_KAPPA=6
run WN_KAPPA=4
run WN_KAPPA=5
run WN_KAPPA=8
run WN_TILE_HEAVY=96
run WN_TILE_HEAVY=384
run WN_BENCH_LEAF=3
run WN_BENCH_LEAF=5
run WN_TILE_BATCH=262144
