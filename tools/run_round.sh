#!/bin/bash
# Round evidence pass on a GPU box: GPU test suite, smoke, the bench line, the reference arm, and the ncu launch list of
# the bench command (taken after the same command has exited 0 without ncu).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02}
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference_line.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --batches-per-step 1 --no-decode --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_bench_launches.csv $CMD > gpurun_out/${tag}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${tag}_bench_launches.csv 40 > gpurun_out/${tag}_bench_launches.txt
python - <<PY
import json
l = json.load(open("gpurun_out/${tag}_bench_line.json"))
print("value", l["value"], "ms/batch", l["config"]["ms_per_batch"], "e2e", l["e2e"]["value"], "frac", l["roofline"]["frac"])
print("cpu", l["cpu_baseline"]["value"], "torch_gpu", l.get("torch_gpu_baseline", {}).get("value"))
print("edgeconv", json.dumps(l["edgeconv"]["three_layers_fwd_bwd_ms"]))
print("decode", {k: v["ms"] for k, v in l["decode"].items() if isinstance(v, dict)})
r = json.load(open("gpurun_out/${tag}_bench_reference_line.json"))
print("reference", r["value"], r["cpu_baseline"]["sample"][:80])
PY
