#!/bin/bash
# One GPU-box pass: parity tests, the headline bench, the decode sweep.  Output under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench rc=$?"
cut -c1-900 gpurun_out/bench_train.json
: > gpurun_out/bench_decode.jsonl
for n in 1024 2048 4096 8192; do
  for m in single complete; do
    extra=""; [ "$n" != 1024 ] && extra="--no-cpu-baseline"
    timeout 300 python bench.py --workload decode --decode-n $n --method $m --steps 3 --warmup 3 $extra >> gpurun_out/bench_decode.jsonl 2>> gpurun_out/bench_decode.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_decode.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'], 'ms', d['ms_per_step'], 'val', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'clk', d['clocks']['sm_mhz'], d.get('cpu_baseline', {}).get('value'))
PY
