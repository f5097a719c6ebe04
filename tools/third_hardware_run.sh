#!/usr/bin/env bash
# Round 2, third N=1 call: the complete GPU suite exactly as the driver runs it, SpMM variants incl. the L2-hint one,
# LightGCN with it, then the ncu evidence (launch lists + one --set full capture per hot kernel).
set -u
out=gpurun_out/third_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 1500 python -m pytest tests/ -x -q -m gpu > "$out/pytest_gpu_all.log" 2>&1; echo "pytest -m gpu (all): exit $? -- $(tail -1 "$out/pytest_gpu_all.log")"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke: exit $? -- $(tail -1 "$out/smoke.log")"
timeout 600 python tools/bench_graph.py --spmm-only > "$out/bench_spmm.jsonl" 2> "$out/bench_spmm.err"; echo "bench_spmm: exit $?"; grep -h "variant\|rowsplit_f32" "$out/bench_spmm.jsonl" | cut -c1-160
for v in 0 6; do for b in 1 4; do
  QREC_SPMM_VARIANT=$v QREC_LGCN_ITEM_BLOCKS=$b timeout 600 python bench.py --steps 3 --warmup 3 --no-parity --no-neumf --no-extras --no-roofs > "$out/bench_lgcn_v${v}_b$b.json" 2> "$out/bench_lgcn_v${v}_b$b.err"
  python -c "import json; d=json.load(open('$out/bench_lgcn_v${v}_b$b.json')); lg=d['lightgcn']; print('lightgcn variant $v blocks $b:', {k:round(v['ms_per_step'],3) for k,v in lg.items() if k.startswith('batch')}, 'spmm', lg.get('spmm',{}).get('ms'))"
done; done
NCU="ncu --clock-control none"
for t in neumf lightgcn; do
  timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file "$out/launches_r2_$t.csv" python tools/ncu_targets.py $t > "$out/ncu_launch_$t.log" 2>&1; echo "launch list $t: exit $?"
done
bash tools/profile_r2.sh > "$out/profile_r2.log" 2>&1; echo "profile_r2: exit $?"; tail -12 "$out/profile_r2.log"
