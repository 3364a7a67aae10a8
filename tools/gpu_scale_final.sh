# final scaling refresh: N=8 (under gpurun --gpus 8) or N=1
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-8}; TAG=r2g
run() { name=$1; shift
  if [ "$N" = 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err; fi
  echo "== $name N=$N rc=$?"; tail -1 gpurun_out/${TAG}_${name}_n${N}.json | cut -c1-200; }
run c2 --no-extras
if [ "$N" != 1 ]; then MTTS_DP_OVERLAP=1 run c2ov --no-extras; fi
run c3w --workload c3 --decode-weak
run c5 --workload c5 --steps 2 --warmup 1
