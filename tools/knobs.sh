P='import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["value"],3), round(d["ms_per_step"],3), round(d["e2e"]["value"],3))'
run() { echo -n "$* : "; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "$P"; }
run WN_TILE_BATCH=131072
run WN_TILE_BATCH=32768
run WN_TILE_BATCH=16384
run WN_TILE_BATCH=8192
run WN_TILE_BATCH=4096
