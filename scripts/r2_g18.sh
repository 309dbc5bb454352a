#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2_g18_pytest.txt 2>&1
grep -v "^$\|^user\|^sys" gpurun_out/r2_g18_pytest.txt | tail -30
