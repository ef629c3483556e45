mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --batch 128 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
python -c "import json;d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]);print('2 GPUs:', d['value'], d['e2e']['value'], d['n_gpus'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tools/tiled_run.py 2160 3840 3 > gpurun_out/tiled_2gpu.json 2> gpurun_out/tiled_2gpu.err
tail -1 gpurun_out/tiled_2gpu.json | cut -c1-400
