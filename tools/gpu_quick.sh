mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k2_stride2 or short_batch or golden_fixtures" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_quick.log
{
for shape in "64 64 64 16 32" "32 32 32 32 64" "16 16 16 64 128" "8 8 8 128 256"; do
  timeout 60 build/kbench k2w 2 $shape 2 10 | grep KBENCH
done
timeout 60 build/kbench wgrad 2 16 16 16 128 128 2 10 | grep KBENCH
} > gpurun_out/kbench_quick.txt 2>&1
cat gpurun_out/kbench_quick.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
