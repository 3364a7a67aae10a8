cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -q --timeout 120 2>&1 | tail -15 > gpurun_out/r2k_gemm_tests.log
timeout 600 python tools/bench_gemm.py gpurun_out/r2k_gemm_bench.json > gpurun_out/r2k_gemm_bench.log 2>&1
tail -3 gpurun_out/r2k_gemm_tests.log; cat gpurun_out/r2k_gemm_bench.log | tail -20
