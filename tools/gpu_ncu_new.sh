# ncu --set full captures of the round-2 kernels (final binary): one forward + the first backward launches of the 2^3 kernels, the deep-level filter gradient
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k2_tc_kernel|k2_wgrad_tc_kernel" -c 10 -f -o gpurun_out/prof_k2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad5_deep_kernel -c 4 -f -o gpurun_out/prof_deep python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_deep.log 2>&1
for f in prof_k2 prof_deep; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
python tools/ncu_summary.py gpurun_out/prof_k2.raw.csv gpurun_out/prof_deep.raw.csv > gpurun_out/ncu_full_summary_new_kernels.txt
rm -f gpurun_out/*.ncu-rep
grep -c "^---" gpurun_out/ncu_full_summary_new_kernels.txt
