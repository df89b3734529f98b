# Round evidence on one B200 (gpurun, one GPU): bench lines of both arms, per-layer table, ncu launch list and --set full
# captures, kernel micro-benchmarks, compute-sanitizer passes.  Outputs land in gpurun_out/; copy what is judged to profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --per-layer gpurun_out/per_layer_bf16x3.json > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
kill $SMI
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline --per-layer gpurun_out/per_layer_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python bench.py --config 3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config3_bf16.json 2> gpurun_out/bench_config3.err
python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config5_1gpu_bf16.json 2> gpurun_out/bench_config5.err
python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_refpath.json 2> gpurun_out/bench_fp32.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv5_col_kernel -c 2 -f -o gpurun_out/prof_col python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_col.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv5_tc_kernel -c 2 -f -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:wgrad5_tc_kernel -c 3 -f -o gpurun_out/prof_wgrad python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_wgrad.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"bn_apply_v4|bn_bwd_reduce_v4|bn_bwd_apply_v4|bn_stats_v4|k2_gather_mma|k2_scatter_mma" -s 4 -c 8 -f -o gpurun_out/prof_hbm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_hbm.log 2>&1
for f in prof_col prof_conv prof_wgrad prof_hbm; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
python tools/ncu_summary.py gpurun_out/prof_col.raw.csv gpurun_out/prof_conv.raw.csv gpurun_out/prof_wgrad.raw.csv gpurun_out/prof_hbm.raw.csv > gpurun_out/ncu_full_summary.txt
rm -f gpurun_out/*.ncu-rep
{
echo "# tools/kbench (CUDA events, 5 launches after 3 warm-up launches; VNB_KB_DBG=1 MMA-warp counters of CTA 0 / last CTA); prec 2 = bf16x3, 1 = bf16"
export VNB_KB_DBG=1
build/kbench fprop 2 128 128 128 16 16 2 5
VNB_TC_NO_COL=1 build/kbench fprop 2 128 128 128 16 16 2 5
build/kbench fprop 2 128 128 128 16 16 1 5
build/kbench fprop 2 128 128 128 16 16 2 5 5 16
build/kbench fprop 2 128 128 128 16 32 2 5
build/kbench fprop 2 64 64 64 32 32 2 5
build/kbench fprop 2 64 64 64 32 32 1 5
build/kbench fprop 2 64 64 64 32 32 2 5 5 32
build/kbench fprop 2 32 32 32 64 64 2 5
build/kbench fprop 2 16 16 16 128 128 2 5
build/kbench fprop 2 8 8 8 256 256 2 5
build/kbench wgrad 2 128 128 128 16 16 2 5
build/kbench wgrad 2 128 128 128 16 16 2 5 5 16
build/kbench wgrad 2 64 64 64 32 32 2 5
VNB_WG_NCH1=1 build/kbench wgrad 2 64 64 64 32 32 2 5
build/kbench wgrad 2 64 64 64 32 32 1 5
build/kbench wgrad 2 64 64 64 32 32 2 5 5 32
build/kbench wgrad 2 32 32 32 64 64 2 5
build/kbench wgrad 2 16 16 16 128 128 2 5
build/kbench wgrad 2 8 8 8 256 256 2 5
build/kbench k2g 2 64 64 64 16 32 5
build/kbench k2s 2 64 64 64 16 32 5
build/kbench k2w 2 64 64 64 16 32 5
} > gpurun_out/kbench.txt 2>&1
bash tools/sanitize_run.sh > gpurun_out/sanitize.log 2>&1
cat gpurun_out/sanitizer_summary.txt
cut -c1-300 gpurun_out/bench_bf16x3.json; cut -c1-200 gpurun_out/bench_reference_arm.json; ls gpurun_out | head -60
