#!/bin/bash
# The round's closing single-GPU measurements (run under gpurun from the repo root); every ncu pass follows a plain run of the same command.
set -u
O=gpurun_out/final_r3
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/tests.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "reference arm rc=$?"
for cfg in elasticity3d newton mk solve irregular; do
  timeout 300 python bench.py --config $cfg > $O/cfg_$cfg.json 2> $O/cfg_$cfg.err; echo "config $cfg rc=$?"
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
timeout 200 $CMD > $O/plain_full.log 2>&1; echo "plain full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/launches_n100.csv $CMD > $O/ncu_full.log 2>&1; echo "ncu list rc=$?"
CMD48="python bench.py --n 48 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
timeout 200 $CMD48 > $O/plain48.log 2>&1; echo "plain48 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bilinear_pts|k_imls_p|k_nb_cluster" -s 9 -c 6 -o $O/prof_r3_final -f $CMD48 > $O/ncu48.log 2>&1; echo "ncu full rc=$?"
ls -la $O
