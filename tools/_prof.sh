set -x
python tools/prof_scan.py > gpurun_out/r2_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"scan_(fwd|bwd)_kernel" -s 2 -c 2 -o gpurun_out/r1f_scan -f python tools/prof_scan.py > gpurun_out/r1f_ncu.log 2>&1
ncu -i gpurun_out/r1f_scan.ncu-rep --page raw --csv > gpurun_out/r1f_scan_raw.csv 2>/dev/null
ncu -i gpurun_out/r1f_scan.ncu-rep --page source --csv > gpurun_out/r1f_scan_src.csv 2>/dev/null
python tools/prof_step.py > gpurun_out/r1f_step_plain.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1f_train_step_launches.csv python tools/prof_step.py > gpurun_out/r1f_step_ncu.log 2>&1
