#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_paged.py -x -q -m gpu -s 2>&1 | grep -E "pinned|passed|failed|Error" | head -8
( time python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench3.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','e2e')})
for k in ('decode','notebook_II','configs_1_2','config_5'):
    print(k, json.dumps(d.get(k))[:1800])
PY
tail -5 gpurun_out/bench3.err
