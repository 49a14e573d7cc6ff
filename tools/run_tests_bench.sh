# usage: bash tools/run_tests_bench.sh <tag>  -- quick guard, the sweep-related GPU tests, then a short bench
tag=$1
timeout 120 python -m pytest tests/test_gpu_rsgm.py -m gpu -x -q -k "compute_rsgm_vs_oracle or sweep_cluster_strips" 2>&1 | tail -15 > gpurun_out/${tag}_quick.log
tail -4 gpurun_out/${tag}_quick.log
if ! grep -q " passed" gpurun_out/${tag}_quick.log || grep -q "failed\|error" gpurun_out/${tag}_quick.log; then echo QUICK_FAILED; cat gpurun_out/${tag}_quick.log; exit 1; fi
timeout 500 python -m pytest tests/test_gpu_rsgm.py tests/test_gpu_benchpath.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-steps 0 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench.err
python tools/show.py gpurun_out/${tag}_bench.json
