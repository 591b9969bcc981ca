set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench8.log 2> gpurun_out/r2_bench8.err; echo "rc=$?" >> gpurun_out/r2_bench8.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2_ncu8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_march_flux|k_gen_faces" -c 5 -o gpurun_out/r2_prof_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fast-mode > gpurun_out/r2_ncu8b.log 2>&1
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench8_ref.log 2> gpurun_out/r2_bench8_ref.err
