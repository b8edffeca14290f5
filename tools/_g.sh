mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "knn" 2>&1 | tail -3
python tools/prof_ops.py --ops knn3 --reps 20 --time
ncu --set full --clock-control none --import-source on -k regex:"knn_d3" -c 1 -o gpurun_out/knn_d3_full python tools/prof_ops.py --ops knn3 --reps 1 --warm 1 > gpurun_out/ncu_knn3.log 2>&1
