python tools/step_probe.py bf16x3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --per-layer gpurun_out/per_layer_bf16x3.json > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; cut -c1-260 gpurun_out/bench_bf16x3.json
