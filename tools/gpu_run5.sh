cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${TAG:-r2v}
python tools/prof_step.py > gpurun_out/prof_step_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_step_launches.csv python tools/prof_step.py > gpurun_out/prof_step_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_train_step_launches.csv 60 > gpurun_out/${TAG}_train_step_launches.txt
cat gpurun_out/${TAG}_train_step_launches.txt
