BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
show() { python - "$1" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("%-30s kernel_ms %.4f step_ms %.4f pcg it/s %.1f" % (sys.argv[1], d["roofline"]["kernel_ms"], d["ms_per_step"], d["pcg"]["iters_per_s"]))
PY
}
cp build/diag/libhcurl_zb8.so /tmp/libzb8.so
HCURL_LIB_PATH=/tmp/libzb8.so $BENCH > gpurun_out/d_zb8.json 2>/dev/null; show gpurun_out/d_zb8.json
HCURL_LIB_PATH=/tmp/libzb8.so $BENCH --order 1 --n 255 > gpurun_out/d_zb8_p1.json 2>/dev/null; show gpurun_out/d_zb8_p1.json
$BENCH > gpurun_out/d_zb4.json 2>/dev/null; show gpurun_out/d_zb4.json
HCURL_PATCH_CTAS_PER_SM=6 $BENCH > gpurun_out/d_zb4_c6.json 2>/dev/null; show gpurun_out/d_zb4_c6.json
HCURL_LIB_PATH=/tmp/libzb8.so ncu --set full --clock-control none --import-source on -k regex:'hex_patch_p2' -s 3 -c 1 -f -o gpurun_out/prof_p2_zb8 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3 > /dev/null 2>&1
