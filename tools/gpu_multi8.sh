# 8-GPU evidence (gpurun --gpus 8): gradient equality per schedule, weak-scaling lines per schedule, configs[3] and configs[4]
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in direct ring; do
  VNB_ALLREDUCE=$m $TR --master-port 29541 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" | sed "s/^/[$m] /" | tee -a gpurun_out/dp_check_$N.txt
done
run() { tag=$1; shift; cfg=$1; shift; env "$@" $TR --master-port 29542 bench.py --gpus $N $cfg --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_${N}gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag: %.1f patches/s  %.2f ms/step  e2e %.1f" % (j["value"], j["ms_per_step"], j["e2e"]["value"]))
except Exception as e:
    print("$tag: failed", e)
PY
}
run direct "" VNB_ALLREDUCE=direct
run ring "" VNB_ALLREDUCE=ring
run nccl "" VNB_ALLREDUCE=nccl
run direct_nodual "" VNB_ALLREDUCE=direct VNB_COMM_DUAL_WAIT=0
run config4_bf16 "--config 4" VNB_ALLREDUCE=direct
run config4_bf16_ring "--config 4" VNB_ALLREDUCE=ring
run config5_bf16 "--config 5" VNB_ALLREDUCE=direct
