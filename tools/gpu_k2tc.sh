# tcgen05 / TMA 2^3 kernels: GPU op tests, single-kernel timings against the mma.sync kernels, quick bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k2_stride2" > gpurun_out/pytest_k2.log 2>&1; echo "pytest k2 exit $?"; tail -5 gpurun_out/pytest_k2.log
{
for shape in "64 64 64 16 32" "32 32 32 32 64" "16 16 16 64 128" "8 8 8 128 256"; do
  for op in k2g k2s; do
    timeout 60 build/kbench $op 2 $shape 2 10
    VNB_K2_NO_TC=1 timeout 60 build/kbench $op 2 $shape 2 10
  done
  timeout 60 build/kbench k2s 2 $shape 2 10 5 1
  timeout 60 build/kbench k2w 2 $shape 2 10
  VNB_K2_NO_TC=1 timeout 60 build/kbench k2w 2 $shape 2 10
done
} > gpurun_out/kbench_k2.txt 2>&1
cat gpurun_out/kbench_k2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradients_match_oracle or three_training_steps or config1_64cube or golden_fixtures or short_batch" > gpurun_out/pytest_net.log 2>&1; echo "pytest net exit $?"; tail -5 gpurun_out/pytest_net.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-400 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
VNB_K2_NO_TC=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick_nok2tc.json 2> gpurun_out/bench_quick_nok2tc.err; cut -c1-400 gpurun_out/bench_quick_nok2tc.json
