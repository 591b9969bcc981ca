mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/mgpu_check.py > gpurun_out/r2_mgpu22.log 2>&1; echo "rc=$?" >> gpurun_out/r2_mgpu22.log
