mkdir -p gpurun_out/ab
L=ab_libs
ULTRASPHERE_B200_LIB=$PWD/$L/v22.so timeout 900 python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -5
export AB_BLOCK=100 AB_SETTLE=200 AB_REPS=3
timeout 300 python tools/ab_inproc.py hopf:3 f64 128000000 $L/v21.so $L/v22.so 2>&1 | tee gpurun_out/ab/atan_hopf.log
timeout 300 python tools/ab_inproc.py standard_prime:8 f64 128000000 $L/v21.so $L/v22.so 2>&1 | tee gpurun_out/ab/atan_stdp8.log
timeout 300 python tools/ab_inproc.py standard:9 f64 128000000 $L/v21.so $L/v22.so 2>&1 | tee gpurun_out/ab/atan_std9.log
timeout 300 python tools/ab_inproc.py random:32:0 f32 64000000 $L/v18.so $L/v21.so 2>&1 | tee gpurun_out/ab/ring_v21.log
