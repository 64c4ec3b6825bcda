mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 0"
ncu --set full --clock-control none --import-source on -k regex:'hex_patch_modal' -s 3 -c 1 -f -o gpurun_out/prof_v3_fe64 $BENCH > gpurun_out/prof_v3_fe64.log 2>&1
HCURL_PATCH_DENSE=1 ncu --set full --clock-control none --import-source on -k regex:'hex_patch_modal' -s 3 -c 1 -f -o gpurun_out/prof_v3_dense $BENCH > gpurun_out/prof_v3_dense.log 2>&1
ls -la gpurun_out | tail -5
