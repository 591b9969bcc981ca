mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rans_gpu.py -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
