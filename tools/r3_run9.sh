python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r3_t9.log 2>&1; tail -5 gpurun_out/r3_t9.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r3_e2e.err | tail -1 > gpurun_out/r3_e2e.json
python -c "
import json
d=json.loads(open('gpurun_out/r3_e2e.json').read()); print(d['config']['phase_ms'], d['ms_per_step'], d['config']['pattern_build_ms'], d['e2e']['ms_per_step'], d['e2e']['step_ms'], d['e2e']['breakdown_ms'], d['parity_check']['ok'], d['parity_check']['K_pattern_bit_exact'])"
