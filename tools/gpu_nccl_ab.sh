cd $GRAFT_REPO_ROOT
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'])"; }
run base
MTTS_DP_BUCKET_MB=512 run one_bucket
NCCL_MIN_NCHANNELS=32 run min32ch
NCCL_MIN_NCHANNELS=32 MTTS_DP_BUCKET_MB=512 run both
