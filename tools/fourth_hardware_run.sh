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
for t in spmm; do
  ncu --clock-control none --set full --import-source on -k regex:spmm_rowsplit -s 1 -c 1 -o "$out/spmm_full_r2" -f python tools/ncu_targets.py spmm 2 > "$out/ncu_spmm.log" 2>&1
  ncu -i "$out/spmm_full_r2.ncu-rep" --page raw --csv > "$out/spmm_full_r2_raw.csv" 2>/dev/null; echo "ncu spmm: $(wc -c < "$out/spmm_full_r2_raw.csv") bytes"
done
