#!/usr/bin/env bash
set -u
out=gpurun_out/fourth_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 1800 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu_all.log" 2>&1; echo "pytest -m gpu (all, no -x): exit $? -- $(tail -1 "$out/pytest_gpu_all.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu_all.log" | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > "$out/bench_full.json" 2> "$out/bench_full.err"; echo "bench: exit $?"
python - <<PY
import json
d=json.loads([l for l in open('$out/bench_full.json') if l.startswith('{')][-1])
pc=d.get('parity_check',{})
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'], 'frac_row_op', d['roofline'].get('row_op_peak',{}).get('frac_of_row_op_ceiling'))
print('parity', pc.get('loss_sum_neg_log_sigmoid',{}).get('rel_err'), pc.get('P',{}).get('rms_err_over_rms_update'), pc.get('Q',{}).get('rms_err_over_rms_update'))
print('lightgcn', {k:round(v['ms_per_step'],3) for k,v in d['lightgcn'].items() if k.startswith('batch')}, d['lightgcn'].get('spmm',{}).get('ms'))
nm=d['neumf']; print('neumf', {k:{p:round(nm[k][p]['ms_per_step'],3) for p in ('gmf','mlp','neumf')} for k in ('reference_2d_5d_2d_d','baseline_256_128_64')})
PY
for b in 8 16; do
  QREC_LGCN_ITEM_BLOCKS=$b timeout 600 python bench.py --steps 3 --warmup 3 --no-parity --no-neumf --no-extras --no-roofs > "$out/bench_lgcn_b$b.json" 2> "$out/bench_lgcn_b$b.err"
  python -c "import json; d=json.loads([l for l in open('$out/bench_lgcn_b$b.json') if l.startswith('{')][-1]); lg=d['lightgcn']; print('lightgcn blocks $b:', {k:round(v['ms_per_step'],3) for k,v in lg.items() if k.startswith('batch')})"
done
python - <<PY
import torch, time
from qrec_b200 import engine as E, synthetic
dev=torch.device('cuda',0)
data=synthetic.make_interactions(65536,100000,50,device=dev); P,Q=synthetic.init_tables(65536,100000,64,seed=1,device=dev)
users=torch.arange(65536,dtype=torch.int32,device=dev)
for N in (10,100):
    E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(3): E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N)
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/3
    print('K8 score_topn 65536 users x 100K items N=%d: %.2f ms = %.1f TFLOP/s, %.2f M users/s' % (N, ms, 65536*1e5*128/ms/1e9, 65536/ms/1e3))
PY
for t in spmm; do
  ncu --clock-control none --set full --import-source on -k regex:spmm_rowsplit -s 1 -c 1 -o "$out/spmm_full_r2" -f python tools/ncu_targets.py spmm 2 > "$out/ncu_spmm.log" 2>&1
  ncu -i "$out/spmm_full_r2.ncu-rep" --page raw --csv > "$out/spmm_full_r2_raw.csv" 2>/dev/null; echo "ncu spmm: $(wc -c < "$out/spmm_full_r2_raw.csv") bytes"
done
