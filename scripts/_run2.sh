mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 0"
ncu --set full --clock-control none --import-source on -k regex:'hex_patch_modal' -s 3 -c 1 -f -o gpurun_out/prof_v4_fe32 $BENCH > gpurun_out/prof_v4_fe32.log 2>&1
timeout 120 python scripts/phase_profile.py 128 2>&1 | tail -9
