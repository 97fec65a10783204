# ncu captures for the pipe-rate evidence (run on the GPU box through gpurun); summaries are made on the
# box, the .ncu-rep files are dropped when they would not fit the 64 MiB return limit.
mkdir -p gpurun_out/r2
OUT=gpurun_out/r2
ncu --query-metrics 2>/dev/null | grep -i "inst_executed_pipe\|pipe_.*cycles_active" | awk '{print $1}' | sort -u > $OUT/metrics_pipe.txt
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_xu.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
P=67108864
for w in "eval --order 4" "eval --order 6" "eval --order 8" "evalf32" "evalc4" "evalall --order 4" "evalall --order 8" "knot" "probefma"; do
  tag=$(echo $w | tr -d ' -')
  ncu --metrics $M --clock-control none -k regex:'k_spline_cell|k_evaluate_all|k_find_interval|k_probe_fma' -c 2 --csv --log-file $OUT/pipe_$tag.csv python tools/prof.py $w --points $P --reps 1 > $OUT/pipe_$tag.log 2>&1
done
for k in 4 8; do
  ncu --set full --clock-control none --import-source on -k regex:k_evaluate_all -c 1 -f -o /tmp/prof_evalall$k python tools/prof.py evalall --order $k --points $P --reps 1 > $OUT/prof_evalall$k.log 2>&1
  python tools/ncu_summary.py /tmp/prof_evalall$k.ncu-rep > $OUT/prof_evalall${k}_summary.txt 2>&1
  python tools/ncu_hot.py /tmp/prof_evalall$k.ncu-rep > $OUT/prof_evalall${k}_hot_sass.txt 2>&1
  ls -la /tmp/prof_evalall$k.ncu-rep >> $OUT/sizes.txt
done
du -sh $OUT
