# deep-level filter-gradient kernel + 2^3 kernels: GPU op tests, single-kernel timings, network parity, quick bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k2_stride2 or conv5_ops_match_torch or short_batch" > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops exit $?"; tail -5 gpurun_out/pytest_ops.log
{
for args in "2 16 16 16 128 128 2 10" "2 16 16 16 128 128 2 10 5 128" "2 8 8 8 256 256 2 10" "2 16 16 16 128 128 1 10" "2 8 8 8 256 256 1 10"; do
  timeout 60 build/kbench wgrad $args
  VNB_WG_NO_DEEP=1 timeout 60 build/kbench wgrad $args | grep KBENCH
done
VNB_WD_KSPLIT=1 timeout 60 build/kbench wgrad 2 16 16 16 128 128 2 10
VNB_WD_KSPLIT=4 timeout 60 build/kbench wgrad 2 16 16 16 128 128 2 10
VNB_WD_KSPLIT=2 timeout 60 build/kbench wgrad 2 8 8 8 256 256 2 10
} > gpurun_out/kbench_deep.txt 2>&1
cat gpurun_out/kbench_deep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gradients_match_oracle or three_training_steps or config1_64cube or golden_fixtures or benchmarked_config2" > gpurun_out/pytest_net.log 2>&1; echo "pytest net exit $?"; tail -5 gpurun_out/pytest_net.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-300 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
VNB_WG_NO_DEEP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick_nodeep.json 2> gpurun_out/bench_quick_nodeep.err; cut -c1-300 gpurun_out/bench_quick_nodeep.json
