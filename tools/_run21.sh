mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mgpu_gpu.py -m gpu -q > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
