python bench.py --n 48 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3_plain48.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_imls_p" -s 3 -c 1 -o gpurun_out/prof_r3_imls_p -f python bench.py --n 48 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/r3_ncu48.log 2>&1
tail -2 gpurun_out/r3_ncu48.log
