# usage: N=2 TAG=r2s WL="c2 c3 c3w c5" bash tools/gpu_scale.sh   (under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${N:-2}; TAG=${TAG:-r2s}; WL=${WL:-"c2 c3 c3w c5"}
run() {  # name, args...
  name=$1; shift
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n${N}.json 2> gpurun_out/${TAG}_${name}_n${N}.err
  fi
  echo "== $name N=$N rc=$?"; tail -1 gpurun_out/${TAG}_${name}_n${N}.json | cut -c1-420; tail -2 gpurun_out/${TAG}_${name}_n${N}.err | cut -c1-300
}
for w in $WL; do
  case $w in
    c2) run c2 --no-extras ;;
    c3) run c3 --workload c3 ;;
    c3w) run c3w --workload c3 --decode-weak ;;
    c5) run c5 --workload c5 --steps 2 --warmup 1 ;;
  esac
done
