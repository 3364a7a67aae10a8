cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/prof_gemm.py > gpurun_out/prof_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 4 -c 4 -o gpurun_out/r2_gemm_prof python tools/prof_gemm.py > gpurun_out/prof_gemm_ncu.log 2>&1
tail -3 gpurun_out/prof_gemm_plain.log gpurun_out/prof_gemm_ncu.log
ls -la gpurun_out/*.ncu-rep
