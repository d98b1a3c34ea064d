mkdir -p gpurun_out/r4
timeout 800 python -m pytest tests -x -q -m gpu > gpurun_out/r4/t1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r4/t1.log
timeout 300 python bench.py > gpurun_out/r4/bench1.json 2> gpurun_out/r4/bench1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4/bench1.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['config'].get('phase_ms'), d['e2e'], d.get('parity_check',{}).get('ok'), d['roofline'].get('frac'), d['roofline'].get('fp64_frac'))
PY
