#!/usr/bin/env bash
# round 2, second session, second single-GPU call: whole GPU suite after the split-row launches, drop-in model step times
# (LightGCN one launch vs one per half, SimGCL, NGCF, NeuMF), full bench line
set -u
out=gpurun_out/r2s2b
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 900 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu_all.log" 2>&1; echo "pytest -m gpu: exit $? -- $(tail -1 "$out/pytest_gpu_all.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu_all.log" | head -12
timeout 600 python tools/bench_models.py --steps 5 > "$out/bench_models.jsonl" 2> "$out/bench_models.err"; echo "bench_models: exit $?"; cut -c1-300 "$out/bench_models.jsonl"; tail -3 "$out/bench_models.err"
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
