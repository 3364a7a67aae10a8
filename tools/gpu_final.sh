cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r2z}
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/${TAG}_tests.log 2>&1
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${TAG}_smoke.log 2>&1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/${TAG}_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/${TAG}_bench.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_smoke.log; cut -c1-300 gpurun_out/${TAG}_bench_ref.log | tail -5; grep -v "^$" gpurun_out/${TAG}_bench.log | tail -4 | cut -c1-400
