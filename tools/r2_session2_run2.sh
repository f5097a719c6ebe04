#!/usr/bin/env bash
# round 2, second session, second single-GPU call: whole GPU suite after the split-row launches, drop-in model step times
# (LightGCN one launch vs one per half, SimGCL, NGCF, NeuMF), the launch list of one sharded LightGCN step, full bench line
set -u
out=gpurun_out/r2s2b
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 900 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu_all.log" 2>&1; echo "pytest -m gpu: exit $? -- $(tail -1 "$out/pytest_gpu_all.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu_all.log" | head -12
timeout 600 python tools/bench_models.py --steps 5 > "$out/bench_models.jsonl" 2> "$out/bench_models.err"; echo "bench_models: exit $?"; cut -c1-300 "$out/bench_models.jsonl"; tail -3 "$out/bench_models.err"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file "$out/launches_lgcn_step.csv" python tools/dist_lightgcn.py --skip-parity --profile-range --steps 2 > "$out/ncu_lgcn.log" 2>&1; echo "ncu lgcn: exit $? $(wc -l < "$out/launches_lgcn_step.csv") lines"
python - <<PY
import csv, collections
try:
    rows=[r for r in csv.reader(l for l in open('$out/launches_lgcn_step.csv') if l.startswith('"'))]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
    agg=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[vi].replace(',','')); v = v/1e3 if r[ui] in ('ns','nsecond') else v
        k=r[ki][:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    print('launch list of 2 steps: %d launches, %.1f us' % (sum(a[0] for a in agg.values()), tot))
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]: print('  %-70s x%-3d %9.1f us  %.1f%%' % (k,a[0],a[1],100*a[1]/tot))
except Exception as e:
    print('launch list parse failed', e)
PY
timeout 1200 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench: exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_n1.json') if l.startswith('{')][-1])
    lg=d.get('lightgcn') or {}
    print('bench: value %.3e e2e %.3e ms/step %.3f k1 %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac']))
    for k,v in lg.items():
        if k.startswith('batch'): print('   lightgcn', k, 'chosen %.3f ms | graph' % v['ms_per_step'], v['ms_per_step_graph_replay'], '| eager %.3f' % v['ms_per_step_eager_launches'], v['cuda_graph'], v['graph_error'])
    print('   spmm', (lg.get('spmm') or {}).get('ms'), json.dumps(lg.get('spmm_halves')), '| lightgcn error:', lg.get('error'))
    print('   neumf', json.dumps(d.get('neumf'))[:400])
except Exception as e:
    print('bench FAILED', e); print(open('$out/bench_n1.err').read()[-1500:])
PY
