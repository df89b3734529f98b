mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k2_stride2 or short_batch or golden_fixtures or gradients_match_oracle or config5_192" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_quick.log
{
for shape in "64 64 64 16 32" "32 32 32 32 64" "16 16 16 64 128" "8 8 8 128 256"; do
  for op in k2g k2s; do timeout 60 build/kbench $op 2 $shape 2 10 | grep KBENCH; done
  timeout 60 build/kbench k2s 2 $shape 2 10 5 1 | grep KBENCH
done
} > gpurun_out/kbench_quick.txt 2>&1
cat gpurun_out/kbench_quick.txt
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print(d['value'], d['e2e']['value'], d['parity'])"; tail -3 gpurun_out/bench_quick.err
