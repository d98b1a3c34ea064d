mkdir -p gpurun_out/r4
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r4/launches_e2e.csv $CMD > gpurun_out/r4/ncu_e2e.log 2>&1; echo "ncu list rc=$?"
tail -2 gpurun_out/r4/ncu_e2e.log | cut -c1-300
