mkdir -p gpurun_out/ab
L=ab_libs
ULTRASPHERE_B200_LIB=$PWD/$L/v23.so timeout 900 python -m pytest tests/test_gpu_math.py tests/test_gpu_parity.py -x -q -m gpu -k "f64 or math or sincos or golden or full_size" 2>&1 | tail -3
export AB_BLOCK=100 AB_SETTLE=200 AB_REPS=3
timeout 300 python tools/ab_inproc.py hopf:3 f64 128000000 $L/v22.so $L/v23.so 2>&1 | tee gpurun_out/ab/sincos_hopf.log
timeout 300 python tools/ab_inproc.py standard_prime:8 f64 128000000 $L/v22.so $L/v23.so 2>&1 | tee gpurun_out/ab/sincos_stdp8.log
