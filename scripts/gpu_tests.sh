#!/bin/bash
# GPU parity tests only (all of them, no -x), log to gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
