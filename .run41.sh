mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r41_tests.log 2>&1; echo "rc $?" >> gpurun_out/r41_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r41_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r41_smoke.log
timeout 400 python bench.py > gpurun_out/r41_bench.log 2>&1; echo "rc $?" >> gpurun_out/r41_bench.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r41_launches.csv python bench.py --steps 1 --warmup 0 --seconds-per-gpu 30 --no-cpu-baseline --batch 8 > gpurun_out/r41_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows_kernel -s 40 -c 8 -o gpurun_out/r41_rows_full python bench.py --steps 1 --warmup 0 --seconds-per-gpu 30 --no-cpu-baseline --batch 8 > gpurun_out/r41_ncu2.log 2>&1
for f in tests smoke; do tail -n 3 gpurun_out/r41_$f.log; done; tail -n 2 gpurun_out/r41_bench.log
