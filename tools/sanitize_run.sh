# compute-sanitizer passes over the per-op and small whole-network GPU tests (run under gpurun, one GPU):
#   memcheck  - out-of-bounds / misaligned accesses of every kernel the tests launch
#   racecheck - shared-memory hazards (the mbarrier / TMA pipelines are outside its model; it covers the epilogues,
#               the 2x2x2 tiles and the reduction kernels)
#   synccheck - divergent barriers
# The sanitizer slows kernels by 10-100x: only the small shapes are selected, each tool under its own timeout.
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer_summary.txt
SEL='conv5_ops_match_torch and bf16x3 or conv3_ops_match_torch and bf16x3 and dims0 or golden_fixtures and tiny or short_batch or k2_stride2 and dims0 or bn_softmax'
RACE_SEL='conv5_ops_match_torch and bf16x3 and 16-16-dims0 or k2_stride2 and 16-32-dims0 and bf16x3'   # racecheck is ~10x slower again
for tool in memcheck racecheck synccheck; do
  sel="$SEL"; [ $tool = racecheck ] && sel="$RACE_SEL"
  timeout 360 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_$tool.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > gpurun_out/sanitizer_$tool.pytest.log 2>&1
  echo "$tool exit $?" | tee -a gpurun_out/sanitizer_summary.txt
  grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitizer_$tool.log | sed "s/^/$tool clean processes: /" | tee -a gpurun_out/sanitizer_summary.txt
done
