mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke20.log 2>&1; echo "rc=$?" >> gpurun_out/r2_smoke20.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench20.log 2> gpurun_out/r2_bench20.err; echo "rc=$?" >> gpurun_out/r2_bench20.err
timeout 900 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench20_c5.log 2> gpurun_out/r2_bench20_c5.err; echo "rc=$?" >> gpurun_out/r2_bench20_c5.err
