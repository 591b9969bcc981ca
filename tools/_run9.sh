set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo8.txt 2>&1; nproc >> gpurun_out/r2_topo8.txt; free -g >> gpurun_out/r2_topo8.txt
timeout 1200 python -m pytest tests/test_mgpu_gpu.py -m gpu -q > gpurun_out/r2_pytest9_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench9_n8.log 2> gpurun_out/r2_bench9_n8.err; echo "rc=$?" >> gpurun_out/r2_bench9_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench9_n4.log 2> gpurun_out/r2_bench9_n4.err; echo "rc=$?" >> gpurun_out/r2_bench9_n4.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 > gpurun_out/r2_bench9_c5_n8.log 2> gpurun_out/r2_bench9_c5_n8.err; echo "rc=$?" >> gpurun_out/r2_bench9_c5_n8.err
