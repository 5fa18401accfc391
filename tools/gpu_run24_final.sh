#!/bin/bash
# final check of the round: full GPU suite, smoke(), default bench line, reference arm
set -x
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02ac_pytest.log
tail -4 gpurun_out/r02ac_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02ac_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02ac_smoke.log
tail -2 gpurun_out/r02ac_smoke.log
timeout 900 python bench.py > gpurun_out/r02ac_bench_n1.json 2> gpurun_out/r02ac_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ac_ref_n1.json 2> gpurun_out/r02ac_ref_n1.err; echo "ref rc=$?"
