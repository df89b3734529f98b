mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q -x --durations=8) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -16 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
