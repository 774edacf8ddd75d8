#!/bin/bash
# Rebuild with different bin sizes (GS_BIN_TILES x 16 pixels) ON THE GPU BOX and time config 2 (and optionally others).
#   tools/bin_sweep.sh 4 6 8
mkdir -p gpurun_out
for bt in "$@"; do
  GS_NVCC_EXTRA="-DGS_BIN_TILES=$bt" python -c "
import importlib
gs = importlib.import_module('aframe-gaussian-splatting_b200')
gs.build.build_library(force=True)
" || exit 1
  timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_dist.py -m gpu -q -x 2>&1 | tail -1
  timeout 300 python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/binsweep_$bt.json 2> gpurun_out/binsweep_$bt.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/binsweep_$bt.json").read().strip().splitlines()[0])
print("BIN_TILES=$bt", "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"], 4) for k, v in d["stages"].items()}, d["counters"]["n_instances_kept"], d["gpu_launches"])
PY
done
