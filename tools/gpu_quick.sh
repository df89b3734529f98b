mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixtures or gradients_match_oracle or three_training_steps or bn_softmax or short_batch or dropout" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_quick.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl_$i.json 2> gpurun_out/bench_pdl_$i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl_$i.json')); print('pdl   ', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
VNB_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopdl_$i.json 2> gpurun_out/bench_nopdl_$i.err; python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl_$i.json')); print('no pdl', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl_bf16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_pdl_bf16.json')); print('pdl bf16', d['value'], d['ms_per_step'], d['e2e']['value'])"
VNB_NO_PDL=1 python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_nopdl_bf16.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_nopdl_bf16.json')); print('no pdl bf16', d['value'], d['ms_per_step'], d['e2e']['value'])"
