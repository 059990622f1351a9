mkdir -p gpurun_out
set -x
timeout 120 python tests/diag_fused.py > gpurun_out/r39_diag.log 2>&1; echo "diag rc $?" >> gpurun_out/r39_diag.log
timeout 400 python -m pytest tests/test_gpu_tc_conv.py -x -q > gpurun_out/r39_tc.log 2>&1
VR_FUSE_UP=1 timeout 500 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r39_par.log 2>&1
for f in 0 1; do VR_FUSE_UP=$f timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r39_bench$f.log 2>&1; done
for f in diag tc par; do tail -n 3 gpurun_out/r39_$f.log; done
grep -o '"ms_per_step": [0-9.]*\|"achieved": [0-9.]*' gpurun_out/r39_bench0.log gpurun_out/r39_bench1.log
