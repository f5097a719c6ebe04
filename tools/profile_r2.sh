#!/usr/bin/env bash
# Round-2 ncu evidence (one gpurun call, 1 GPU): the launch list of a bench run, then one --set full capture per
# hot kernel.  Results: gpurun_out/r2/*.ncu-rep + *.csv; summarise here with tools/ncu_summary.py into profiles/.
set -u
out=gpurun_out/r2
mkdir -p "$out"
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$out/launches_r2_bench.csv" \
  python bench.py --steps 2 --warmup 1 --no-parity --no-lightgcn --no-neumf --no-extras --no-roofs > "$out/bench_under_ncu.log" 2>&1
for t in ${TARGETS:-k1_sig k1 k1_hbm rowops spmm topn}; do
  kre="regex:bpr_sgd_usermajor"
  [ "$t" = rowops ] && kre="regex:row_op_kernel"
  [ "$t" = spmm ] && kre="regex:spmm_"
  [ "$t" = topn ] && kre="regex:score_topn"
  skip=1; count=1; reps=2
  [ "$t" = rowops ] && { skip=0; count=3; reps=1; }          # the three modes, one launch each
  timeout 900 $NCU --set full --import-source on -k "$kre" -s $skip -c $count -o "$out/${t}_full_r2" -f \
    python tools/ncu_targets.py "$t" $reps > "$out/ncu_$t.log" 2>&1
  echo "$t: exit $?"
  ncu -i "$out/${t}_full_r2.ncu-rep" --page raw --csv > "$out/${t}_full_r2_raw.csv" 2>/dev/null
done
ls -la "$out"
