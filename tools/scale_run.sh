python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2v_n1.json 2>gpurun_out/r2v_n1.err
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2v_n$n.json 2>gpurun_out/r2v_n$n.err; done
for n in 1 2 4 8; do python -c "
import json
l=json.loads(open('gpurun_out/r2v_n$n.json').read().strip().splitlines()[-1])
print('N=$n', round(l['value'],3), round(l['e2e']['value'],3), round(l['ms_per_step'],4), l['inside_count'], l['parity']['strict_band_mismatches'], l['clocks']['sm_mhz'], l['clocks']['reasons'], l['gpu_launches'], l.get('tree_broadcast_ms_steady'))"; done
