mkdir -p gpurun_out
export VNB_KB_DBG=1
(
build/kbench fprop 1 16 16 16 32 32 2 2 || exit 1
python -m pytest tests/test_gpu_parity.py -x -q -k "conv5_ops or conv3_ops" 2>&1 | tail -3
for tm in 1 2; do
export VNB_TC_TMAX=$tm
echo "=== TMAX $tm"
build/kbench fprop 2 64 64 64 32 32 2 5
build/kbench fprop 2 64 64 64 32 32 1 5
build/kbench fprop 2 64 64 64 64 32 2 5
build/kbench fprop 2 32 32 32 64 64 2 5
build/kbench fprop 2 32 32 32 64 64 1 5
build/kbench fprop 2 16 16 16 128 128 2 5
build/kbench fprop 2 16 16 16 128 128 1 5
done
unset VNB_TC_TMAX
build/kbench fprop 2 128 128 128 16 16 2 5
) > gpurun_out/kb1.log 2>&1
grep -v "timed out" gpurun_out/kb1.log | head -90
