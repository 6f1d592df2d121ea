O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_gpu_tests.log 2>&1; tail -2 $O/r02_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err; tail -1 $O/bench_final.err
