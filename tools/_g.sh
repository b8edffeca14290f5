mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "sampler or loss" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train3.json 2> gpurun_out/bench_train3.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_train3.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_train3.json').read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e_host_sampled_triplets'], 'losses', d['loss'], d['e2e_loss'], d['e2e_loss_device_sampler'], 'launches', d['gpu_launches'])
PY
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import bench, hpcs_b200 as hb
dev = torch.device('cuda:0')
host = bench.synth_inputs(32, 0)
order, seg, T0 = hb.triplet_plan(host['labels'], 50, 0.0)
plan = (order.to(dev), seg.to(dev), T0)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
print('device sampler us (kernel)', t(lambda: hb.sample_triplets_device(None, seed=1, plan=plan)), 'T0', T0)
PY
