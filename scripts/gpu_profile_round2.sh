#!/bin/bash
# round-2 call: t5 rev2 (lean epilogue), grouped k_dw_tc, cooperative k_knn -- tests, timings, ncu
O=gpurun_out/r2j; mkdir -p $O
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/pytest.txt
timeout 300 python scripts/time_t5.py > $O/time_t5.txt 2>&1
QT_PROFILE=1 timeout 600 python scripts/quick_time_all.py > $O/qt.json 2> $O/qt_prof.txt
for c in vox point nice coslam; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
done
XRD_PREC=1 timeout 300 $NCU --set full --import-source on -k regex:k_fused -s 2 -c 1 -o $O/coslam_g2 python scripts/ncu_coslam.py 4096 > $O/ncu_coslam.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:'k_dw_tc|k_gemm_t5' -s 8 -c 4 -o $O/vox python bench.py --config vox --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_vox.log 2>&1
timeout 400 $NCU --set full -k regex:'k_knn|k_nb_build_bwd' -s 3 -c 2 -o $O/point python bench.py --config point --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_point.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/launches_vox.csv python bench.py --config vox --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_l_vox.log 2>&1
tail -8 $O/pytest.txt; cat $O/time_t5.txt; cat $O/qt.json
for c in coslam vox point nice; do python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$c.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('$c', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', r.get('frac'), r.get('kernel_ms'))
except Exception as e:
    print('$c ERR', e); print(open('$O/bench_$c.err').read()[-800:])
PY
done
ls -la $O
