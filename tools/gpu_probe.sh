mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=8 -k "dropout_with or loss_zoo or optimizers_follow or sync_bn_over") > gpurun_out/pytest_gpu.log 2>&1
tail -16 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_report.txt
