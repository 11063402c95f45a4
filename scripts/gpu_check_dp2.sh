#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
timeout 600 python -m pytest tests/test_dp_gpu.py -q 2>&1 | tail -15 > $O/dp_pytest.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 --no-cpu-baseline --no-torch-gpu-baseline > $O/bench_coslam_n2.json 2> $O/bench_coslam_n2.err
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 10 --scaling strong --no-cpu-baseline --no-torch-gpu-baseline > $O/bench_coslam_n2_strong.json 2> $O/bench_coslam_n2_strong.err
timeout 600 $TR bench.py --config nice --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_nice_n2.json 2> $O/bench_nice_n2.err
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > $O/bench_coslam_n1.json 2> $O/bench_coslam_n1.err
cat $O/dp_pytest.txt; for f in coslam_n1 coslam_n2 coslam_n2_strong nice_n2; do echo "== $f"; python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$f.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','dp_parity')}, d['e2e']['value'], d.get('torch_gpu_baseline'))
except Exception as e:
    print('ERR', e); print(open('$O/bench_$f.err').read()[-1500:])
PY
done
