mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_mgpu_gpu.py -m gpu -q > gpurun_out/r2_pytest15_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest15_mgpu.log
for n in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench15_n$n.log 2> gpurun_out/r2_bench15_n$n.err; echo "rc=$?" >> gpurun_out/r2_bench15_n$n.err
done
