#!/bin/bash
# One GPU-box call that produces the round's evidence under gpurun_out/ (tools/make_profiles_r02.py
# turns it into profiles/r02_*):   gpurun --timeout 2400 -- 'bash tools/evidence_run.sh r02'
R=${1:-r02}
O=gpurun_out
mkdir -p $O
# 1. bounds-asserting build under the smoke of every kernel family (compute-sanitizer is closed on this pool)
ULTRASPHERE_B200_LIB=$PWD/ultrasphere_b200/lib/libultrasphere_b200_checked.so timeout 600 python tools/sanitize_smoke.py > $O/${R}_checked_smoke.log 2>&1; echo "checked smoke rc=$?" | tee -a $O/${R}_checked_smoke.log
# 2. the GPU test suite (writes gpurun_out/parity_report.json)
timeout 1500 python -m pytest tests -x -q -m gpu > $O/${R}_gpu_tests.log 2>&1; tail -2 $O/${R}_gpu_tests.log
# 3. both bench arms, the driver's flags
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_${R}_reference.json 2> $O/bench_${R}_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_${R}.json 2> $O/bench_${R}.err; tail -1 $O/bench_${R}.err
# 4. launch list of the bench command (shares, not absolutes)
SHORT="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-quadrature --no-configs --no-sustained"
$SHORT > $O/plain_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_${R}.csv $SHORT > $O/ncu_launches.log 2>&1
# 5. ncu --set full of the hot kernels at the BASELINE sizes
cap() {  # name, regex, skip, count, command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o $O/prof_${R}_$name "$@" > $O/ncu_$name.log 2>&1
  tail -1 $O/ncu_$name.log
}
cap c2 cartesian_tiled 4 2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-quadrature --no-configs --no-sustained
cap hopf3_f64 cartesian_tiled 4 2 python tools/profile_one.py hopf:3 f64 128000000
cap stdp8_f64 cartesian_tiled 4 2 python tools/profile_one.py standard_prime:8 f64 128000000
cap ring cartesian_ring 4 2 python tools/profile_one.py random:32:0 f32 64000000
cap c1 cartesian_tiled 4 2 python tools/profile_one.py spherical f64 1000000
cap c5 "to_cartesian_dot_rows|wreduce_tma|to_cartesian_bcast" 6 3 python tools/profile_c5.py
# 6. per-call host cost, drift of the headline kernel
python tools/host_overhead.py > $O/${R}_host_overhead.json 2>&1
timeout 200 python tools/jitter.py > $O/${R}_jitter.json 2>&1
# 7. clocks, power and throttle reasons under ~2 s of back-to-back launches (the 1000 W cap)
python tools/clocks_probe.py standard:9 f32 256000000 2 > $O/${R}_clocks_std9_f32.json 2>> $O/${R}_clocks.err
python tools/clocks_probe.py hopf:3 f64 128000000 2 > $O/${R}_clocks_hopf3_f64.json 2>> $O/${R}_clocks.err
python tools/clocks_probe.py random:32:0 f32 64000000 2 > $O/${R}_clocks_ring_f32.json 2>> $O/${R}_clocks.err
# 8. compute-sanitizer, if this pool lets it run (it has been refused so far: the log says which)
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $O/${R}_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/${R}_memcheck.log; tail -3 $O/${R}_memcheck.log
ls -la $O | tail -40
