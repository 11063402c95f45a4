#!/bin/bash
# round-2 call: full suite, t5 rev1/rev2 A/B, dW tensor-core kernel, ncu captures, all bench configs
O=gpurun_out/r2i; mkdir -p $O
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > $O/pytest.txt
timeout 300 python scripts/time_t5.py > $O/time_t5.txt 2>&1
timeout 300 python scripts/quick_time_coslam.py 0 -1 -2 > $O/qt_coslam.txt 2>&1
for c in coslam vox point nice; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_$c.json 2> $O/bench_$c.err
done
# ---- ncu: full captures of the dominant kernels
T5_NCU=1 timeout 300 $NCU --set full --import-source on -k regex:k_gemm_t5 -c 4 -o $O/t5 python scripts/time_t5.py > $O/ncu_t5.log 2>&1
XRD_PREC=1 timeout 300 $NCU --set full --import-source on -k regex:k_fused -s 3 -c 1 -o $O/coslam_g2 python scripts/ncu_coslam.py 4096 > $O/ncu_coslam.log 2>&1
timeout 400 $NCU --set full -k regex:'k_dw_tc|k_knn|k_decoder_fwd|k_decoder_bwd|k_nb_build_bwd' -s 4 -c 10 -o $O/point python bench.py --config point --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_point.log 2>&1
# ---- ncu: launch lists of one bench run per config
for c in vox point nice; do
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_l_$c.log 2>&1
done
tail -8 $O/pytest.txt; cat $O/time_t5.txt $O/qt_coslam.txt
for c in coslam vox point nice; do python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$c.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('$c', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', r.get('frac'), r.get('kernel_ms'), d.get('torch_gpu_baseline'))
except Exception as e:
    print('$c ERR', e); print(open('$O/bench_$c.err').read()[-800:])
PY
done
ls -la $O
