cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r2q}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --no-extras > gpurun_out/${TAG}_bench.log 2>&1
timeout 900 python bench.py --workload c3 --decode-steps 512 > gpurun_out/${TAG}_bench_c3.log 2>&1
tail -6 gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-300; tail -3 gpurun_out/${TAG}_bench_c3.log | cut -c1-600
