set -x
T=${TAG:-r1n}
timeout 900 python -m pytest tests/test_encoder_gpu.py -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/${T}_pytest_encoder.log
tail -8 gpurun_out/${T}_pytest_encoder.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
python -c "
import json; j=json.load(open('gpurun_out/${T}_bench_cfg2.json')); print(json.dumps(j['extra'].get('encoder'), indent=1)); print(j['value'], j['e2e'], j['roofline'])"
tail -5 gpurun_out/${T}_bench_cfg2.err
