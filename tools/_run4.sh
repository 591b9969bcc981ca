set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench4_n2.log 2> gpurun_out/r2_bench4_n2.err; echo "rc=$?" >> gpurun_out/r2_bench4_n2.err
