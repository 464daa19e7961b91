python tools/prof_run.py 10000 2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'acc_local|acc_final' -s 2 -c 2 -f -o gpurun_out/prof_acc_r1g python tools/prof_run.py 10000 2 > gpurun_out/ncu_acc.log 2>&1
echo rc=$?
