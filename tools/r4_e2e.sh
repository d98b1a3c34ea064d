mkdir -p gpurun_out/r4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r4/e2e.json 2> gpurun_out/r4/e2e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4/e2e.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['config']['phase_ms'], d['e2e']['ms_per_step'], d['e2e']['step_ms'], d['e2e']['breakdown_ms'], d['parity_check']['ok'])
PY
