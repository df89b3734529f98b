# last pass with the final binary: the tests touching the 2^3 / deep kernels, both bench lines, per-layer tables, launch list
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k2_stride2 or short_batch or golden_fixtures or gradients_match_oracle or three_training_steps or conv5_ops_match_torch or config1_64cube" > gpurun_out/pytest_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_final.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --per-layer gpurun_out/per_layer_bf16x3.json > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
kill $SMI
python -c "
import json; d=json.load(open('gpurun_out/bench_bf16x3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['parity']); print(d['roofline']['achieved'], d['roofline']['wgrad_tflops'])"
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline --per-layer gpurun_out/per_layer_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
cut -c1-200 gpurun_out/bench_bf16.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
head -24 gpurun_out/launch_summary.txt
{
for shape in "64 64 64 16 32" "32 32 32 32 64" "16 16 16 64 128" "8 8 8 128 256"; do
  for op in k2g k2s k2w; do timeout 60 build/kbench $op 2 $shape 2 10 | grep KBENCH; done
  timeout 60 build/kbench k2s 2 $shape 2 10 5 1 | grep KBENCH
done
} > gpurun_out/kbench_k2_final.txt 2>&1
cat gpurun_out/kbench_k2_final.txt
