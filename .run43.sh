mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r43_launches.csv python bench.py --steps 1 --warmup 0 --seconds-per-gpu 30 --no-cpu-baseline --batch 8 > gpurun_out/r43_ncu1.log 2>&1
echo done
