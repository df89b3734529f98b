# 2-GPU sanity line with the final binary (PDL launches + gradient exchange)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_final.json 2> gpurun_out/bench_2gpu_final.err
cut -c1-260 gpurun_out/bench_2gpu_final.json; tail -2 gpurun_out/bench_2gpu_final.err
