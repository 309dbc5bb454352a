#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_small.py > gpurun_out/r2_g31_memcheck.txt 2>&1; tail -4 gpurun_out/r2_g31_memcheck.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2_g31_tests.txt 2>&1; tail -3 gpurun_out/r2_g31_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
