mkdir -p gpurun_out/r4
CMD="python bench.py --n 64 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cluster_build_masks|k_cluster_tables|k_pattern_clusters|k_cluster_pmask" -c 4 -o gpurun_out/r4/prof_index -f $CMD > gpurun_out/r4/ncu_index.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r4/ncu_index.log | cut -c1-200
