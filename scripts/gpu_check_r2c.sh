#!/bin/bash
# round-2 GPU check: new tcgen05 GEMM, the full GPU suite, Co-SLAM kernel A/B + ncu, all bench configs
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python -m pytest tests/test_gemm_t5_gpu.py -q 2>&1 | tail -25 > $O/t5.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest.txt
XRD_PREC=1 timeout 300 python scripts/quick_time_coslam.py 0 -1 2>&1 | tail -12 > $O/coslam_ab.txt
XRD_PREC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 2 -c 1 -o $O/coslam_g1 python scripts/ncu_coslam.py 4096 > $O/ncu.log 2>&1
for c in coslam nice vox point; do
  timeout 900 python bench.py --config $c --steps 40 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err
done
QT_PROFILE=1 timeout 600 python scripts/quick_time_all.py > $O/qt.json 2> $O/qt_prof.txt
head -c 600 $O/t5.txt; echo; tail -5 $O/pytest.txt; cat $O/coslam_ab.txt; for c in coslam nice vox point; do head -c 400 $O/bench_$c.json; echo; tail -3 $O/bench_$c.err; done
