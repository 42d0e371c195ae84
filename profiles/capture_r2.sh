#!/bin/bash
# Round-2 evidence capture (run under gpurun from the repo root). Each ncu pass runs only after the same command has
# exited 0 without ncu; numbers printed by runs under ncu are never bench values.
set -u
OUT=gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --c5-pairs 2048"
$BENCH > $OUT/r2_bench_plain.json 2> $OUT/r2_bench_plain.err || exit 1
# 1. launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches_bench.csv $BENCH > $OUT/r2_bench_under_ncu.log 2>&1
# 2. full capture of the dominant kernel inside the bench command (one launch of k_align_warp, 65,536 hypotheses)
ncu --set full --clock-control none --import-source on -k regex:k_align_warp -s 3 -c 1 -f -o $OUT/r2_full_k_align_warp $BENCH > $OUT/r2_full_warp.log 2>&1
# 3. C5 pairs matcher
python profiles/prof_c5.py 8192 > $OUT/r2_c5_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_align_pairs -s 1 -c 1 -f -o $OUT/r2_full_k_align_pairs python profiles/prof_c5.py 8192 > $OUT/r2_full_pairs.log 2>&1
# 4. grid build at C3 size (4 M points, 0.1 m cells) and C2 size: launch list + full capture of the build kernels
python profiles/prof_c2c3.py > $OUT/r2_c2c3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r2_launches_c2c3.csv python profiles/prof_c2c3.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"k_count|k_alloc|k_fill|k_rank|k_finalize|k_bounds" -s 14 -c 6 -f -o $OUT/r2_full_grid_c3 python profiles/prof_c2c3.py > $OUT/r2_full_grid.log 2>&1
ls -la $OUT | grep r2_
# 5. small batch: team of warps per match (k_align_team)
python profiles/prof_team.py 4096 > $OUT/r2_team_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_align_team -s 3 -c 1 -f -o $OUT/r2_full_k_align_team python profiles/prof_team.py 4096 > $OUT/r2_full_team.log 2>&1
