cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q --timeout 120 -x 2>&1 | tail -15 > gpurun_out/r2o_gemm_tests.log
timeout 600 python tools/bench_gemm.py gpurun_out/r2o_gemm_bench.json > gpurun_out/r2o_gemm_bench.log 2>&1
tail -3 gpurun_out/r2o_gemm_tests.log; cut -c1-150 gpurun_out/r2o_gemm_bench.log | tail -22
