set -x
T=${TAG:-r1k}
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -q --timeout 120 -x 2>&1 | tail -25 > gpurun_out/${T}_pytest_search.log
tail -5 gpurun_out/${T}_pytest_search.log
grep -q " passed" gpurun_out/${T}_pytest_search.log || exit 1
grep -q " failed" gpurun_out/${T}_pytest_search.log && exit 1
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
cat gpurun_out/${T}_bench_cfg2.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_launches_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:merge_rescore -s 3 -c 1 -o gpurun_out/${T}_prof_merge_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_merge_cfg2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 1 -o gpurun_out/${T}_prof_scan_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_scan_cfg2.log 2>&1
