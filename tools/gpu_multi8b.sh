# 8-GPU lines with the final round-2 kernels (gpurun --gpus 8): default schedule, configs[1] bf16x3 and configs[3] bf16
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VNB_ALLREDUCE=direct $TR --master-port 29541 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" | tee gpurun_out/dp_check_8b.txt
$TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_8gpu_final.json 2> gpurun_out/bench_8gpu_final.err
cut -c1-260 gpurun_out/bench_8gpu_final.json; tail -2 gpurun_out/bench_8gpu_final.err
$TR --master-port 29543 bench.py --gpus $N --config 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_8gpu_config4_final.json 2> gpurun_out/bench_8gpu_config4_final.err
cut -c1-260 gpurun_out/bench_8gpu_config4_final.json
