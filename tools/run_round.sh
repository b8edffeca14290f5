#!/bin/bash
# One GPU-box pass producing the round's evidence: parity tests, bench lines, ncu launch lists and full captures.
# Usage (from the repo root, through gpurun):  bash tools/run_round.sh <tag>      e.g. r01_c
# Every ncu command runs only after the same command has exited 0 without ncu.  Output: gpurun_out/<tag>_*
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -2 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
# ---- headline bench
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_line.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_line.json 2>> $out/${tag}_bench.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
# ---- decode lines (configs[4]) and launch lists
: > $out/${tag}_decode_lines.jsonl
for m in single complete; do for n in 1024 2048 4096 8192; do
  extra="--no-cpu-baseline"; [ "$n" = 1024 ] && extra=""
  timeout 600 python bench.py --workload decode --decode-n $n --method $m --steps 3 --warmup 3 $extra >> $out/${tag}_decode_lines.jsonl 2>> $out/${tag}_bench.err
done; done
for n in 1024 8192; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_decode_launches_$n.csv \
      python bench.py --workload decode --decode-n $n --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
done
# ---- full captures of the hot kernels (one launch each; source lines imported)
python tools/prof_ops.py --ops knn3,knn63,edge,loss,sampler --reps 1 --warm 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on \
    -k regex:"knn_d3|knn_tc_kernel|knn_rerank32|knn_pack|edge_feat_fwd|edge_rev2|edge_bwd_gather|hyp_triplet_kernel|hyp_bwd|triplet_sample" \
    -c 14 -o $out/${tag}_ops_full python tools/prof_ops.py --ops knn3,knn63,edge,loss,sampler --reps 1 --warm 0 > $out/${tag}_ncu_ops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pdist|linkage_kernel|boruvka" -c 9 -o $out/${tag}_decode_full \
    python bench.py --workload decode --decode-n 2048 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_decode.log 2>&1
ls -la $out | grep ${tag}
