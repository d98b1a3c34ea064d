#!/bin/bash
# Assembly-kernel cost breakdown: bench phases with parts of k_bilinear_reg disabled (FG_DEBUG_SKIP; results are wrong by design).
for m in ${MODES:-0 1 4 3}; do
  FG_DEBUG_SKIP=$m python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > /tmp/asm_mode.json
  python - "$m" <<'PY'
import json, sys
d = json.loads(open("/tmp/asm_mode.json").read())
print("mode", sys.argv[1], d["config"].get("phase_ms"))
PY
done
