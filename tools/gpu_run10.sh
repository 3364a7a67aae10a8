cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r2y}
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
timeout 600 python tools/bench_gemm.py gpurun_out/${TAG}_gemm_bench.json > gpurun_out/${TAG}_gemm_bench.log 2>&1
timeout 600 python tools/prof_ops.py gpurun_out/${TAG}_prof_ops.json > gpurun_out/${TAG}_prof_ops.log 2>&1
timeout 900 python bench.py --no-extras > gpurun_out/${TAG}_bench.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log; cut -c1-140 gpurun_out/${TAG}_gemm_bench.log | tail -19; head -30 gpurun_out/${TAG}_prof_ops.log; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-300
for m in 0 1; do MTTS_DP_OVERLAP=$m timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras 2>/dev/null | cut -c1-220; done
