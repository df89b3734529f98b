mkdir -p gpurun_out
export VNB_KB_DBG=1
(
build/kbench wgrad 1 16 16 16 16 16 2 2 || exit 1
python -m pytest tests/test_gpu_parity.py -x -q -k "conv5_ops or conv3_ops" 2>&1 | tail -3
build/kbench wgrad 2 128 128 128 16 16 2 5
build/kbench wgrad 2 128 128 128 16 16 1 5
build/kbench wgrad 2 64 64 64 32 32 2 5
build/kbench wgrad 2 32 32 32 64 64 2 5
build/kbench wgrad 2 16 16 16 128 128 2 5
build/kbench wgrad 2 8 8 8 256 256 2 5
) > gpurun_out/kb1.log 2>&1
grep -v "timed out" gpurun_out/kb1.log | head -90
