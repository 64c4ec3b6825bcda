mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pcg.py tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --pcg-iters 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pcg', d['pcg'], 'step %.4f'%d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 -k regex:'pcg_|hex_patch_modal|dot' --csv --log-file gpurun_out/l_D.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 4 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/l_D.csv')) if len(r)>10 and r[0].isdigit()]
print([(r[4][:24], r[-1]) for r in rows][-8:])
PY
