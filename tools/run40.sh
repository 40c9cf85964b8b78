python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/b40_2gpu.json 2> gpurun_out/b40_2gpu.err
tail -c 1800 gpurun_out/b40_2gpu.json; tail -3 gpurun_out/b40_2gpu.err
