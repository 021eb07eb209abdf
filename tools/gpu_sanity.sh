#!/bin/bash
# last check of a committed state: smoke + the force / PME / variant parity tests (about 30 s)
OUT=gpurun_out/sanity; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forces_match or pme or variants or laplacian" > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
