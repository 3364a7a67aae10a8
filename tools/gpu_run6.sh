cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -15 > gpurun_out/r2m_tests.log
timeout 600 python tools/prof_ops.py gpurun_out/r2m_prof_ops.json > gpurun_out/r2m_prof_ops.log 2>&1
timeout 900 python bench.py > gpurun_out/r2m_bench.log 2>&1
tail -3 gpurun_out/r2m_tests.log; cat gpurun_out/r2m_prof_ops.log; tail -c 3000 gpurun_out/r2m_bench.log
