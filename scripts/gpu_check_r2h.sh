#!/bin/bash
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python -m pytest tests/test_coslam_gpu.py tests/test_bench_shapes_gpu.py tests/test_algorithms_gpu.py -m gpu -q -x 2>&1 | tail -15 > $O/pytest.txt
timeout 200 python scripts/quick_time_coslam.py 0 -1 > $O/qt_jac.txt 2>&1
XRD_COSLAM_NO_JAC=1 timeout 200 python scripts/quick_time_coslam.py 0 > $O/qt_nojac.txt 2>&1
timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_coslam.json 2> $O/bench_coslam.err
timeout 400 python bench.py --config vox --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_vox.json 2> $O/bench_vox.err
cat $O/pytest.txt $O/qt_jac.txt $O/qt_nojac.txt; head -c 1500 $O/bench_coslam.json; echo; tail -3 $O/bench_coslam.err; head -c 900 $O/bench_vox.json; tail -3 $O/bench_vox.err
