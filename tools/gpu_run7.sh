cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/prof_gemm2.py > gpurun_out/prof_gemm2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 5 -c 5 -o gpurun_out/r2m_gemm_epi python tools/prof_gemm2.py > gpurun_out/prof_gemm2_ncu.log 2>&1
tail -3 gpurun_out/prof_gemm2_plain.log gpurun_out/prof_gemm2_ncu.log
