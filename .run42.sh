mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/r42_tests.log 2>&1; echo "rc $?" >> gpurun_out/r42_tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r42_smoke.log 2>&1; echo "rc $?" >> gpurun_out/r42_smoke.log
timeout 300 python bench.py > gpurun_out/r42_bench.log 2>&1; echo "rc $?" >> gpurun_out/r42_bench.log
for f in tests smoke; do tail -n 3 gpurun_out/r42_$f.log; done; tail -n 2 gpurun_out/r42_bench.log | cut -c 1-400
