set -x
timeout 1500 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 900 -x 2>&1 | tail -15
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extra | cut -c1-1400
