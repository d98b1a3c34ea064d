mkdir -p gpurun_out/ab
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/ab/bench_n2.json 2> gpurun_out/ab/bench_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/ab/bench_n2.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['config']['phase_ms'], d['e2e']['ms_per_step'], d.get('parity_check'))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/ab/bench_n2.err').read()[-1500:])
PY
