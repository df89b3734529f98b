python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rP -k "sync_bn_over" 2>&1 | grep -E "SYNC_BN_CHECK|passed|failed" | tee gpurun_out/sync_bn_2gpu.txt
