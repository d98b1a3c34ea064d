python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r3_t4.log 2>&1; tail -3 gpurun_out/r3_t4.log
for v in base nopf; do
  O=""
  if [ $v == nopf ]; then O="--opt asm_prefetch=0"; fi
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity $O 2> gpurun_out/r3_$v.err | tail -1 > gpurun_out/r3_$v.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r3_$v.json').read()); print('$v', d['config']['phase_ms'], d['ms_per_step'])"
done
