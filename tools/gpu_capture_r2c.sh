# Round-2 ncu evidence, part c: the solve kernels after the backward blocks were decoupled from the halo.
OUT=gpurun_out/r3
mkdir -p $OUT
set -x
cap() {   # name, kernel regex, prof.py arguments, launches to skip
  python tools/prof.py $3 --reps 1 > $OUT/$1_plain.log 2>&1 || return
  ncu --set full --clock-control none --import-source on -k regex:$2 -s ${4:-0} -c 1 -f -o /tmp/$1 python tools/prof.py $3 --reps 1 > $OUT/$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep > $OUT/$1_summary.txt 2>&1
  python tools/ncu_hot.py /tmp/$1.ncu-rep > $OUT/$1_hot_sass.txt 2>&1
}
cap prof_tmem_r3 k_banded_tmem "solve" 1
cap prof_tmem32_r2 k_banded_tmem "solve32w" 1
cap prof_fused_r2 k_banded_tmem "fused" 1
