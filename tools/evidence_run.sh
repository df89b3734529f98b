# Round evidence on one B200: GPU parity tests, bench lines (both arms), ncu launch list and --set full captures.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 3 --per-layer gpurun_out/per_layer_bf16x3.json > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline --e2e-staged --per-layer gpurun_out/per_layer_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv5_tc_kernel -c 2 -f -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad5_tc_kernel -c 2 -f -o gpurun_out/prof_wgrad python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"bn_apply_v4|bn_bwd_reduce_v4|k2_gather_mma" -s 2 -c 6 -f -o gpurun_out/prof_hbm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_hbm.log 2>&1
for f in prof_conv prof_wgrad prof_hbm; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_bf16x3.json gpurun_out/bench_bf16.json gpurun_out/bench_reference.json; ls -la gpurun_out
