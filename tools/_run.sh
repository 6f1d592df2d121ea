set -x
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02b_gpu_tests.log 2>&1; tail -3 $O/r02b_gpu_tests.log
python tools/profile_one.py random:32:0 f32 64000000 > $O/plain_ring.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cartesian_ring -s 4 -c 2 -f -o $O/prof_r02b_ring python tools/profile_one.py random:32:0 f32 64000000 > $O/ncu_ring.log 2>&1
tail -2 $O/ncu_ring.log
