mkdir -p gpurun_out/r4
CMD="python bench.py --n 64 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_imls_p" -s 1 -c 1 -o gpurun_out/r4/prof_imls -f $CMD > gpurun_out/r4/ncu_imls.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r4/ncu_imls.log | cut -c1-200
