python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu -k "kriging or lid or cavity or mk or MK" > gpurun_out/r3_t5.log 2>&1; tail -5 gpurun_out/r3_t5.log
for v in base old; do
  O=""
  if [ $v == old ]; then O="--opt mk_rows=0"; fi
  python bench.py --config mk --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity $O 2> gpurun_out/r3_mk_$v.err | tail -1 > gpurun_out/r3_mk_$v.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r3_mk_$v.json').read()); print('$v', d['config'].get('phase_ms'), d['ms_per_step'], d['config'].get('checks'))"
done
