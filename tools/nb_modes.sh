#!/bin/bash
# Neighbour-search cluster size sweep (FG_NB_UCAP): bench phases only.
for u in ${UCAPS:-32 64 100 128}; do
  FG_NB_UCAP=$u python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/nb_mode.json
  python - "$u" <<'PY'
import json, sys
d = json.loads(open("gpurun_out/nb_mode.json").read())
print("nb_ucap", sys.argv[1], d["config"].get("phase_ms"), d["value"])
PY
done
