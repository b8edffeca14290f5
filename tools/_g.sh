mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "linkage" 2>&1 | tail -5
: > gpurun_out/bench_decode4.jsonl
for n in 1024 2048 4096 8192; do
    timeout 300 python bench.py --workload decode --decode-n $n --method single --steps 3 --warmup 3 --no-cpu-baseline >> gpurun_out/bench_decode4.jsonl 2>> gpurun_out/bench_decode4.err
done
tail -5 gpurun_out/bench_decode4.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_decode4.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'], 'ms', d['ms_per_step'], 'val', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
PY
for n in 1024 8192; do
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches_$n.csv python bench.py --workload decode --decode-n $n --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - $n <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(f'gpurun_out/decode_launches_{sys.argv[1]}.csv')) if len(r) > 10 and r[0].isdigit()]
seen = {}
for r in rows[-40:]:
    name = r[4].split('(')[0][:40]
    seen.setdefault(name, []).append(float(r[-1]) / 1e3)
for k, v in seen.items():
    print(sys.argv[1], k, [round(x, 1) for x in v[-6:]])
PY
done
ncu --set full --clock-control none --import-source on -k regex:"pdist|linkage_kernel|contract|rowmin" -c 7 -o gpurun_out/decode_full python bench.py --workload decode --decode-n 2048 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out/*.ncu-rep
