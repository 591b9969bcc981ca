mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mgpu_gpu.py -m gpu -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench13_n2.log 2> gpurun_out/r2_bench13_n2.err; echo "rc=$?" >> gpurun_out/r2_bench13_n2.err
