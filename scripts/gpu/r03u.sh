#!/bin/bash
# Round 2, GPU call AU: FINAL validation at HEAD (full GPU suite, smoke, default bench line,
# reference arm).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/parity_r02.jsonl
(timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30) > $O/r03u_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r03u_smoke.log 2>&1
python bench.py > $O/r03u_bench_default.json 2> $O/r03u_bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r03u_bench_reference_arm.json 2>&1
