set -x
T=${TAG:-r1attn}
timeout 1200 python -m pytest tests/test_decoder_gpu.py tests/test_hybrid_gpu.py -m gpu -q --timeout 600 -x 2>&1 | tail -30 > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_pytest.log
grep -q " failed\|rror" gpurun_out/${T}_pytest.log && exit 1
timeout 900 python - <<'PY'
import json, torch, bench
print(json.dumps(bench.decoder_numbers(torch.device("cuda", 0), cpu=False)))
PY
LXG_DECODER_PACK=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/${T}_launches_rerank.csv python scripts/gpu_decoder_prof.py > /dev/null 2>&1
python scripts/ncu_launches.py gpurun_out/${T}_launches_rerank.csv > gpurun_out/${T}_launches_rerank.txt 2>&1; cat gpurun_out/${T}_launches_rerank.txt
