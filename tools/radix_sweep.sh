#!/bin/bash
# Rebuild with different radix chunk sizes (256 x GS_RADIX_ITEMS elements) ON THE GPU BOX and time config 2 (+3, 4).
mkdir -p gpurun_out
for it in "$@"; do
  GS_NVCC_EXTRA="-DGS_RADIX_ITEMS=$it" python -c "
import importlib
gs = importlib.import_module('aframe-gaussian-splatting_b200')
gs.build.build_library(force=True)
" || exit 1
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/radixsweep_$it.json 2> gpurun_out/radixsweep_$it.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/radixsweep_$it.json").read().strip().splitlines()[0])
print("RADIX_ITEMS=$it", "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"], 4) for k, v in d["stages"].items()}, [(o["config"]["workload"][:8], round(o["value"]), {k: round(v["ms"], 3) for k, v in o["stages"].items()}) for o in d.get("other_configs", [])])
PY
done
