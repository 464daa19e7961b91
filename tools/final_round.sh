out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/pytest_r2u.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest_r2u.log
python bench.py --steps 10 --warmup 3 > $out/bench_r2u.json 2> $out/bench_r2u.err; echo "bench rc=$?"
python tools/prof_run.py 10000 2 > $out/plain2_r2u.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'d8_kernel|acc_local|super_tile|coarse_resolve|acc_final' -s 6 -c 6 -f -o $out/prof_path_r2u python tools/prof_run.py 10000 2 > $out/ncu_path_r2u.log 2>&1
echo "path capture rc=$?"
