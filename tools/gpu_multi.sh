# multi-GPU evidence (run with gpurun --gpus N): gradient equality per all-reduce schedule, sync-BN over NCCL, weak-scaling lines
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in direct ring nccl; do
  VNB_ALLREDUCE=$m $TR --master-port 29541 tools/dp_check.py 2>&1 | grep -E "DP_CHECK|Error|error" | sed "s/^/[$m] /" | tee -a gpurun_out/dp_check_$N.txt
done
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rP -k "sync_bn_over" 2>&1 | grep -E "SYNC_BN_CHECK|passed|failed" | tee gpurun_out/sync_bn_2gpu.txt
fi
run() { tag=$1; shift; env "$@" $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$tag.json 2> gpurun_out/bench_${N}gpu_$tag.err; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_${N}gpu_$tag.json").read().strip().splitlines()[-1])
    print("$tag: %.1f patches/s  %.2f ms/step  e2e %.1f" % (j["value"], j["ms_per_step"], j["e2e"]["value"]))
except Exception as e:
    print("$tag: failed", e)
PY
}
run direct VNB_ALLREDUCE=direct
run direct_nodual VNB_ALLREDUCE=direct VNB_COMM_DUAL_WAIT=0
run ring VNB_ALLREDUCE=ring
run nccl VNB_ALLREDUCE=nccl
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu: %.1f patches/s %.2f ms/step' % (j['value'], j['ms_per_step']))"
