#!/bin/bash
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > $O/pytest.txt
for c in coslam nice vox point; do
  timeout 900 python bench.py --config $c --steps 40 --warmup 5 > $O/bench_$c.json 2> $O/bench_$c.err
done
QT_PROFILE=1 timeout 600 python scripts/quick_time_all.py > $O/qt.json 2> $O/qt_prof.txt
tail -30 $O/pytest.txt; for c in coslam nice vox point; do head -c 300 $O/bench_$c.json; echo; tail -2 $O/bench_$c.err; done; cat $O/qt.json
