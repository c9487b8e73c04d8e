#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_paligemma.py tests/test_gpu_kernels.py -x -q -m gpu -k "paligemma or attn_fwd or rope" > gpurun_out/pytest_pg2.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_pg2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-decode --no-slots --no-configs-1-2 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_c5.json') if l.startswith('{')][-1])
print(json.dumps(d.get('config_5'))[:1500])
PY
tail -3 gpurun_out/bench_c5.err
