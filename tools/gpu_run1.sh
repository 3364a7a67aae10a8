set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for sel in "layouts and 0-0" "layouts and 0-1" "layouts and 1-0" "layouts and 1-1" "bias_accumulate" "batched" "weight_gradient" "gelu" "softmax" "persistent"; do
  timeout 300 python -m pytest tests/test_gpu_gemm.py -q -x --timeout 120 -k "$sel" 2>&1 | tail -15
done > gpurun_out/r2a_gemm_tests.log 2>&1
timeout 600 python -m pytest tests/test_reference_pin.py tests/test_gpu_model.py -q -m gpu --timeout 300 -k "reference or context_not" > gpurun_out/r2a_pin_tests.log 2>&1
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu --timeout 120 -k "gemm_bf16 or noncontiguous" > gpurun_out/r2a_ops_tests.log 2>&1
timeout 600 python tools/bench_gemm.py gpurun_out/r2a_gemm_bench.json > gpurun_out/r2a_gemm_bench.log 2>&1
tail -5 gpurun_out/r2a_gemm_tests.log gpurun_out/r2a_pin_tests.log gpurun_out/r2a_ops_tests.log; cat gpurun_out/r2a_gemm_bench.log | tail -20
