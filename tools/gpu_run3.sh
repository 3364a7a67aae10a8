cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_reference_pin.py -q -m gpu --timeout 300 -k "not bf16_c" 2>&1 | tail -30 > gpurun_out/r2e_model_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench.log 2>&1
tail -30 gpurun_out/r2e_model_tests.log; tail -5 gpurun_out/r2e_bench.log | cut -c1-3000
