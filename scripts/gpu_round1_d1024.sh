set -x
T=${TAG:-r1d1024}
timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_decoder_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
LXG_SCAN_ASMEM=0 timeout 900 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 600 -k "768 or ragged or adversarial or golden" 2>&1 | tail -10 > gpurun_out/${T}_pytest_asmem768.log
cat gpurun_out/${T}_pytest_asmem768.log
timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_nt64.json 2> gpurun_out/${T}_bench_cfg3.err
LXG_SCAN_ASMEM=0 timeout 400 python bench.py --workload cfg3 --steps 30 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/${T}_bench_cfg3_asmem4.json 2>> gpurun_out/${T}_bench_cfg3.err
python - <<PY
import json
for f in ("nt64", "asmem4"):
    j = json.load(open("gpurun_out/${T}_bench_cfg3_%s.json" % f)); print(f, j["value"], j["e2e"]["value"], j["roofline"]["ms_per_launch"], j["roofline"]["frac"])
PY
