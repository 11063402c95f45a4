#!/bin/bash
# end-of-round check: what the driver runs (GPU tests, smoke, default bench, reference arm) + all configs
O=gpurun_out/r2k; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/pytest.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $O/smoke.txt 2>&1
( time timeout 900 python bench.py ) > $O/bench_coslam.json 2> $O/bench_coslam.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
for c in nice vox point; do
  ( time timeout 900 python bench.py --config $c ) > $O/bench_$c.json 2> $O/bench_$c.err
done
tail -12 $O/pytest.txt; tail -5 $O/smoke.txt
for c in coslam reference nice vox point; do python - <<PY
import json
try:
    d=json.loads([l for l in open('$O/bench_$c.json').read().strip().splitlines() if l.startswith('{')][-1])
    r=d.get('roofline') or {}
    print('$c', {k:d.get(k) for k in ('value','ms_per_step','steps','gpu_launches')}, 'e2e', d['e2e']['value'], 'roof', r.get('frac'), r.get('kernel_ms'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'tgpu', (d.get('torch_gpu_baseline') or {}).get('value'))
except Exception as e:
    print('$c ERR', e); print(open('$O/bench_$c.err').read()[-1200:])
PY
tail -4 $O/bench_$c.err | grep real
done
