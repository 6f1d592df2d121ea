mkdir -p gpurun_out/ab
L=ab_libs
AB_BLOCK=100 AB_SETTLE=200 AB_REPS=3 timeout 300 python tools/ab_inproc.py random:32:0 f32 64000000 $L/v23.so $L/v25p1.so $L/v25p2.so $L/v25p3.so 2>&1 | tee gpurun_out/ab/ring_policy.log
for v in v25p1 v25p2 v25p3; do
ULTRASPHERE_B200_LIB=$PWD/$L/$v.so ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:from_cartesian_ring -s 2 -c 1 --csv python tools/profile_one.py random:32:0 f32 64000000 2>&1 | grep "dram__bytes_read\|gpu__time" | sed 's/.*Command line profiler metrics",//' | tee -a gpurun_out/ab/ring_policy.log
done
