python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/b40.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('host_affinity'))"
tail -3 gpurun_out/b40.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/b40_2gpu.json 2> gpurun_out/b40_2gpu.err
tail -c 1500 gpurun_out/b40_2gpu.json; tail -5 gpurun_out/b40_2gpu.err
