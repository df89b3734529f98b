# round-2 call 1: GPU parity suite (incl. the new full-size tests), default bench with per-layer table and staged e2e,
# kernel diagnostics (MMA-warp counters) for the Cout = 16 and filter-gradient kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
(time python -m pytest tests -m gpu -q -x --durations=15) > gpurun_out/pytest_gpu.log 2>&1
tail -30 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --per-layer gpurun_out/per_layer_bf16x3.json --e2e-staged > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
cat gpurun_out/bench_bf16x3.json | cut -c1-1500
{
export VNB_KB_DBG=1
echo "== enc1 fprop 16->16 @128^3"; build/kbench fprop 2 128 128 128 16 16 2 5
echo "== same, T=2"; VNB_TC_TMAX=2 build/kbench fprop 2 128 128 128 16 16 2 5
echo "== same, T=1"; VNB_TC_TMAX=1 build/kbench fprop 2 128 128 128 16 16 2 5
echo "== same, bf16"; build/kbench fprop 2 128 128 128 16 16 1 5
echo "== dec1 fprop 16+16->16"; build/kbench fprop 2 128 128 128 16 16 2 5 5 16
echo "== dec1 dgrad 16->32 (CT=32,KC=16)"; build/kbench fprop 2 128 128 128 16 32 2 5
echo "== dec1 dgrad 16->32 (two CT=16 slices)"; VNB_TC_NO_CT32K16=1 build/kbench fprop 2 128 128 128 16 32 2 5
echo "== enc2 fprop 32->32 @64^3"; build/kbench fprop 2 64 64 64 32 32 2 5
echo "== wgrad 16->16 @128^3"; build/kbench wgrad 2 128 128 128 16 16 2 5
echo "== wgrad 32->16 @128^3"; build/kbench wgrad 2 128 128 128 16 16 2 5 5 16
echo "== wgrad 32->32 @64^3"; build/kbench wgrad 2 64 64 64 32 32 2 5
echo "== wgrad 64->32 @64^3"; build/kbench wgrad 2 64 64 64 32 32 2 5 5 32
echo "== wgrad 64->64 @32^3"; build/kbench wgrad 2 32 32 32 64 64 2 5
} > gpurun_out/kbench_call1.txt 2>&1
cat gpurun_out/kbench_call1.txt
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv5_ops_match_torch and bf16x3 and (16-16-dims0 or 32-16-dims1 or 64-32-dims2)" > gpurun_out/sanitizer_memcheck.pytest.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/sanitizer_memcheck.log
