# Round-2 final evidence on one B200 (gpurun, one GPU) after the tcgen05 2^3 kernels and the deep-level filter gradient:
# full GPU test-suite, bench lines of both arms, per-layer table, ncu launch list, ncu --set full captures of the new kernels,
# compute-sanitizer passes over the new kernels.  Outputs land in gpurun_out/; what is judged is copied to profiles/.
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=12) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -22 gpurun_out/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --per-layer gpurun_out/per_layer_bf16x3.json --e2e-staged > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
kill $SMI
cut -c1-600 gpurun_out/bench_bf16x3.json; tail -2 gpurun_out/bench_bf16x3.err
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline --per-layer gpurun_out/per_layer_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
cut -c1-300 gpurun_out/bench_bf16.json
python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config3_bf16.json 2> gpurun_out/bench_config3.err
python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config5_1gpu_bf16.json 2> gpurun_out/bench_config5.err
cut -c1-300 gpurun_out/bench_config3_bf16.json; cut -c1-300 gpurun_out/bench_config5_1gpu_bf16.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference.err
cut -c1-300 gpurun_out/bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
head -40 gpurun_out/launch_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k2_tc_kernel|k2_wgrad_tc_kernel" -c 10 -f -o gpurun_out/prof_k2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad5_deep_kernel -c 4 -f -o gpurun_out/prof_deep python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_deep.log 2>&1
for f in prof_k2 prof_deep; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
python tools/ncu_summary.py gpurun_out/prof_k2.raw.csv gpurun_out/prof_deep.raw.csv > gpurun_out/ncu_full_summary_new_kernels.txt
rm -f gpurun_out/*.ncu-rep
SEL='k2_stride2 and bf16x3 or conv5_ops_match_torch and bf16x3 and (128-128-dims11 or 256-128-dims12 or 128-256-dims13) or short_batch and bf16x3'
for tool in memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_new_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_new_$tool.pytest.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_new_summary.txt
  tail -2 gpurun_out/sanitizer_new_$tool.pytest.log | tee -a gpurun_out/sanitizer_new_summary.txt
  grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitizer_new_$tool.log | sed "s/^/$tool clean processes: /" | tee -a gpurun_out/sanitizer_new_summary.txt
done
ls gpurun_out
