mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "linkage or sampler or edge" 2>&1 | tail -3
for n in 1024 2048; do
python bench.py --workload decode --decode-n $n --method single --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'ms', d['ms_per_step'], 'val', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
done
python tools/prof_ops.py --ops edge --reps 20 --time
