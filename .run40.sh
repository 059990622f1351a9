mkdir -p gpurun_out
for u in 2 3 4; do VR_USLOTS=$u timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r40_u$u.log 2>&1; done
for b in 27 41; do timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 --batch $b > gpurun_out/r40_b$b.log 2>&1; done
VR_FUSE_UP=0 timeout 200 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r40_f0.log 2>&1
grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' gpurun_out/r40_*.log
