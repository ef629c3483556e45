mkdir -p gpurun_out
for mode in "--streams 1" "--streams 2" "--streams 2 --stream-priorities"; do
  timeout 300 python bench.py --batch 128 --steps 2 --warmup 2 --no-cpu-baseline $mode > gpurun_out/bench_streams.json 2> gpurun_out/bench_streams.err
  echo "$mode: $(python -c "import json;d=json.load(open('gpurun_out/bench_streams.json'));print(d['value'], d['e2e']['value'])")"
done
