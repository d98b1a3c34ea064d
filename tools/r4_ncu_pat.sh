mkdir -p gpurun_out/r4
CMD="python bench.py --n 64 --steps 1 --warmup 1 --no-cpu-baseline --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_pattern_clusters|k_linear_pts" -c 2 -o gpurun_out/r4/prof_pat -f $CMD > gpurun_out/r4/ncu_pat.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/r4/ncu_pat.log | cut -c1-200
