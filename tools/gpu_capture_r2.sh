# Round-2 ncu evidence (run on the GPU box through gpurun AFTER the same commands exited 0 without ncu).
# Summaries are made on the box; .ncu-rep files stay there (64 MiB return limit) except the headline one.
OUT=gpurun_out/r3
mkdir -p $OUT
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/cap_bench_plain.json 2> $OUT/cap_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches_r4.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $OUT/cap_bench_ncu.log 2>&1
P=268435456
cap() {   # name, kernel regex, prof.py arguments
  python tools/prof.py $3 --reps 1 > $OUT/$1_plain.log 2>&1 || return
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -f -o /tmp/$1 python tools/prof.py $3 --reps 1 > $OUT/$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $OUT/$1_summary.txt 2>&1
  python tools/ncu_hot.py /tmp/$1.ncu-rep > $OUT/$1_hot_sass.txt 2>&1
}
cap prof_cell_r5 k_spline_cell "eval --points $P"
cp /tmp/prof_cell_r5.ncu-rep $OUT/ 2>/dev/null
cap prof_cellg_r2 k_spline_cell "evalc4 --points $P"
cap prof_knot_r2 k_find_interval "knot --points 134217728"
cap prof_evalall4_r2 k_evaluate_all "evalall --order 4 --points 134217728"
VARIANTS=2 cap prof_gather_r1 k_probe_gather "gather --points $P"
ls -la /tmp/*.ncu-rep > $OUT/sizes.txt
