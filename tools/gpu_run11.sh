cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "scan" 2>&1 | tail -6
python tools/prof_gemm.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 4 -o gpurun_out/r2_gemm_main python tools/prof_gemm.py > gpurun_out/ncu1.log 2>&1
python tools/prof_gemm2.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 5 -o gpurun_out/r2_gemm_epi python tools/prof_gemm2.py > gpurun_out/ncu2.log 2>&1
TAG=r2 bash tools/gpu_run5.sh | head -30
ls -la gpurun_out/*.ncu-rep | tail -3
