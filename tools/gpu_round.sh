#!/bin/bash
# One GPU-box pass (run under gpurun): GPU parity tests, the bench line, the ncu
# launch list of the same bench command and one `--set full` capture of every
# kernel of the path.  Outputs land in gpurun_out/; tools/ncu_summary.py turns
# the reports into the committed summaries under profiles/.
#   usage: bash tools/gpu_round.sh <tag> [notest] [noprof]
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
if [[ " $* " != *" notest "* ]]; then
  python -m pytest tests -m gpu -x -q > $out/pytest_$tag.log 2>&1
  echo "pytest rc=$?"; tail -3 $out/pytest_$tag.log
fi
python bench.py --steps 10 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err
echo "bench rc=$?"; cat $out/bench_$tag.json
if [[ " $* " == *" noprof "* ]]; then exit 0; fi
# launch list of the bench command (per-launch times are cold-cache: compare shares)
python bench.py --steps 2 --warmup 1 > $out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    -k regex:'d8_kernel|acc_|coarse_resolve|super_tile|outer_nodes|virtual_nodes|edge_roots|marked_nodes|band_|mark_edges|add_imports|weight_max|terrain_kernel|filter_kernel|run_filter|row_prefix|tile_phase|alt_kernel|fill_|classify_|collect_|ccl_|generation_|quorum_|round_kernel|finish_kernel|deflate_|tile_|pack_i8|synth_' \
    --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 > $out/ncu_bench_$tag.log 2>&1
echo "launch list rc=$?"
# full capture: the four kernels of one D8 + accumulation step (second step: warm)
python tools/prof_run.py 10000 2 > $out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'d8_kernel|acc_local|super_tile|coarse_resolve|acc_final' -s 6 -c 6 -f \
    -o $out/prof_path_$tag python tools/prof_run.py 10000 2 > $out/ncu_path_$tag.log 2>&1
echo "path capture rc=$?"
# the weighted mode's kernels (stage A / B / C in exact fixed point)
python tools/weighted_phases.py > $out/plain4_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'weight_max|acc_local_w|acc_final_w|coarse_resolve_kernel<.*i128' -c 8 -f \
    -o $out/prof_weighted_$tag env HT_PROF=1 python tools/weighted_phases.py > $out/ncu_weighted_$tag.log 2>&1
echo "weighted capture rc=$?"
python tools/prof_terrain.py 16384 0 > $out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:terrain_kernel -s 2 -c 1 -f \
    -o $out/prof_terrain_$tag python tools/prof_terrain.py 16384 0 > $out/ncu_terrain_$tag.log 2>&1
echo "terrain capture rc=$?"
