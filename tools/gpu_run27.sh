#!/bin/bash
# round-2 GPU session 27: tile shape of the residual-add GEMMs in situ at the 8-GPU per-GPU shape (M = 2048):
# default heuristic (192-column tiles for O-proj and FFN-down) against 256-column tiles (--gemm-variant 2), interleaved
set -x
cd "$GRAFT_REPO_ROOT"
A="--query-tokens 4 --skip-secondary --skip-cpu-baseline --parity-queries 0 --skip-e2e --steps 30"
for i in 1 2 3; do
timeout 200 python bench.py $A > gpurun_out/r02ag_auto_$i.json 2> gpurun_out/r02ag_auto_$i.err
timeout 200 python bench.py $A --gemm-variant 2 > gpurun_out/r02ag_v2_$i.json 2> gpurun_out/r02ag_v2_$i.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02ag_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["value"]), round(j["ms_per_step"], 3), j["clocks"]["sm_mhz"], round(j["phases_ms_per_step"]["encode_gemm_ms"], 2))
    except Exception as e:
        print(f, "ERR", e)
PY
