cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r2p}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 600 python tools/prof_ops.py gpurun_out/${TAG}_prof_ops.json > gpurun_out/${TAG}_prof_ops.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log; head -45 gpurun_out/${TAG}_prof_ops.log; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400
