#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 150 $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --no-torch-gpu-baseline > $O/bench_coslam_n2.json 2> $O/bench_coslam_n2.err ) 2> $O/time_coslam_n2.txt
( time timeout 200 $TR bench.py --config nice --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_nice_n2.json 2> $O/bench_nice_n2.err ) 2> $O/time_nice_n2.txt
timeout 150 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -6 > $O/dp_pytest.txt
cat $O/dp_pytest.txt; cat $O/time_coslam_n2.txt $O/time_nice_n2.txt; for f in coslam_n2 nice_n2; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','dp_parity','gpu_launches')}, d['e2e']['value'])
except Exception as e:
    print('ERR', e); print(open('$O/bench_$f.err').read()[-1200:])
PY
done
