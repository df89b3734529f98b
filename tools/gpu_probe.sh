mkdir -p gpurun_out; rm -f gpurun_out/parity_report.txt
(time python -m pytest tests -m gpu -q -x --durations=5) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
cat gpurun_out/parity_report.txt
VNB_NO_FUSED_STATS=1 python tools/step_probe.py bf16x3
python tools/step_probe.py bf16x3
