python tools/filter_bench.py 8192 > gpurun_out/plain_f.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:filter_kernel -s 3 -c 1 -f -o gpurun_out/prof_filter_max python tools/filter_bench.py 8192 > gpurun_out/ncu_f1.log 2>&1
echo rc=$?
