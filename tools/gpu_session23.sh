#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "one_billion" > gpurun_out/pytest_1e9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_1e9.log
tail -25 gpurun_out/pytest_1e9.log
