#!/bin/bash
# quick iteration: selected op tests + model parity + bench. usage: gpu_iter.sh "<pytest -k expr for ops>"
mkdir -p gpurun_out
K="${1:-flash or gemm or conv3x3}"
timeout -k 10 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -x -k "$K" > gpurun_out/iter_ops.log 2>&1
echo "ops exit $?"; tail -n 4 gpurun_out/iter_ops.log
timeout -k 10 600 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -s > gpurun_out/iter_model.log 2>&1
echo "model exit $?"; grep -E "passed|failed|: \{" gpurun_out/iter_model.log | tail -12
timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/iter_bench.json 2> gpurun_out/iter_bench.err
echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/iter_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clocks', d['clocks'])
for k,v in d['kernels'].items(): print(f"  {k:20s} {v['ms_per_step']:8.3f} ms share {v['share']:.3f} ", {kk:round(vv,1) for kk,vv in v.items() if kk in ('tflops','gbs')})
PY
