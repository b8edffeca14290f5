mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
  python - $n <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/scale_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_host', d['e2e_host_sampled_triplets']['value'])
except Exception as e:
    print('N', sys.argv[1], 'failed', e); print(open(f'gpurun_out/scale_{sys.argv[1]}.err').read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload decode --decode-n 1024 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
