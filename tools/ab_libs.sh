# A/B of builds of the library on the headline step (FLEXIGAL_GPU_LIB override): tools/ab_libs.sh libflexigal_gpu.so libflexigal_gpu_variant.so
# (a variant = one translation unit recompiled with a -D switch and linked with the other objects into flexigal_b200/lib/)
mkdir -p gpurun_out/ab
for lib in "$@"; do
  FLEXIGAL_GPU_LIB=$PWD/flexigal_b200/lib/$lib timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab/ab_$lib.json 2> gpurun_out/ab/ab_$lib.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab/ab_$lib.json').read().strip().splitlines()[-1])
    print('$lib', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['config']['phase_ms'].items()}, d.get('parity_check',{}).get('ok'), d.get('parity_check',{}).get('phi_err_vs_truth'), d.get('parity_check',{}).get('dphi_err_vs_truth'))
except Exception as e:
    print('$lib', 'FAILED', e); print(open('gpurun_out/ab/ab_$lib.err').read()[-800:])
PY
done
