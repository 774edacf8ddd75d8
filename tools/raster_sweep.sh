#!/bin/bash
# Rebuild the library with different raster occupancy knobs ON THE GPU BOX and time config 2 with each.
#   tools/raster_sweep.sh "8 4 2" "10 3 4"      (triples: GS_RASTER_MINB GS_RASTER_STAGES GS_RASTER_UNROLL)
mkdir -p gpurun_out
for v in "$@"; do
  set -- $v
  GS_NVCC_EXTRA="-DGS_RASTER_MINB=$1 -DGS_RASTER_STAGES=$2 -DGS_RASTER_UNROLL=${3:-2}" python -c "
import importlib
gs = importlib.import_module('aframe-gaussian-splatting_b200')
gs.build.build_library(force=True)
" || exit 1
  python bench.py --steps 20 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/sweep_$1_$2_${3:-2}.json 2> gpurun_out/sweep_$1_$2_${3:-2}.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/sweep_$1_$2_${3:-2}.json").read().strip().splitlines()[0])
print("MINB=$1 STAGES=$2 UNROLL=${3:-2}", "fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v["ms"], 4) for k, v in d["stages"].items()})
PY
done
