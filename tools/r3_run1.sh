python -m pytest tests/test_gpu_round2.py -x -q -m gpu -k "imls_kernel_variants or headline" > gpurun_out/r3_t1.log 2>&1; tail -3 gpurun_out/r3_t1.log
for v in base p15 nopipe; do
  case $v in
    base) L=""; O="";;
    p15) L="FLEXIGAL_GPU_LIB=$PWD/flexigal_b200/lib/var/lib_p15.so"; O="";;
    nopipe) L=""; O="--opt imls_pipe=0";;
  esac
  env $L python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity $O 2> gpurun_out/r3_$v.err | tail -1 > gpurun_out/r3_$v.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r3_$v.json').read()); print('$v', d['config']['phase_ms'], d['ms_per_step'])"
done
