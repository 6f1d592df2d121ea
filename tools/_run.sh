mkdir -p gpurun_out/ab
V=ab_libs/v30.so
AB_REPS=200 timeout 300 python tools/ab_inproc.py spherical f64 1000000 $V $V:tiled_min_tiles=64 $V:tile_groups=3 $V:tile_groups=4 $V:tile_groups=5 $V:tile_groups=6,tile_bufs=8 $V:tile_groups=6,tile_bufs=14 $V:tile_groups=4,tile_bufs=8 2>&1 | tee gpurun_out/ab/c1_scan.log
