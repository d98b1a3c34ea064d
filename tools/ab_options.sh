# A/B of fg_set_option switches on the headline step: tools/ab_options.sh "imls_grid=0" "imls_grid=1" ...
mkdir -p gpurun_out/ab
for o in "$@"; do
  tag=$(echo $o | tr '= ' '__')
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $(for kv in $o; do echo --opt $kv; done) > gpurun_out/ab/ab_$tag.json 2> gpurun_out/ab/ab_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab/ab_$tag.json').read().strip().splitlines()[-1])
    print('$o', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['config']['phase_ms'].items()}, d.get('parity_check',{}).get('ok'))
except Exception as e:
    print('$o', 'FAILED', e); print(open('gpurun_out/ab/ab_$tag.err').read()[-800:])
PY
done
