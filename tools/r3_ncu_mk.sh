ncu --set full --clock-control none --import-source on -k regex:"k_mk_rows" -s 3 -c 1 -o gpurun_out/prof_r3_mk -f python bench.py --config mk --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3_ncu_mk.log 2>&1
tail -2 gpurun_out/r3_ncu_mk.log
