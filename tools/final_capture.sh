#!/bin/bash
# The evidence run behind profiles/r2_* (one B200 through gpurun): GPU tests, smoke, both bench arms, the ncu launch list of the
# default bench command, one ncu --set full capture of the tiled kernels, compute-sanitizer passes.
python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; tail -c 300 gpurun_out/r2_final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_tiled.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile|k_plan" -c 5 -f -o gpurun_out/r2_tiled python tools/profile_tiled.py 512 > gpurun_out/r2_tiled_ncu.log 2>&1; tail -2 gpurun_out/r2_tiled_ncu.log
{
for tool in memcheck racecheck initcheck; do echo "## --tool $tool"; timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_small.py 2>&1 | grep -E "sanitize_small done|SUMMARY|ERROR|hazard" | head -5; echo; done
for tool in memcheck racecheck; do echo "## --tool $tool (WN_TILE_LANES=2, WN_TILE_LANES_MIN=1)"; WN_TILE_LANES=2 WN_TILE_LANES_MIN=1 timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_small.py 2>&1 | grep -E "sanitize_small done|SUMMARY|ERROR|hazard" | head -5; echo; done
} > gpurun_out/r2_sanitizer.txt 2>&1
cat gpurun_out/r2_sanitizer.txt
