#!/usr/bin/env bash
# round 2, second session, single-GPU call: whole GPU suite, smoke, the full bench line (LightGCN step replayed from a
# CUDA graph), K2 decomposed into its bipartite halves with the item side column-blocked (tools/bench_spmm_blocks.py)
set -u
out=gpurun_out/r2s2
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
t0=$(date +%s)
timeout 900 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu_all.log" 2>&1; echo "pytest -m gpu: exit $? -- $(tail -1 "$out/pytest_gpu_all.log")  [$(( $(date +%s) - t0 )) s]"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu_all.log" | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke: exit $? -- $(tail -1 "$out/smoke.log")"
t0=$(date +%s)
timeout 900 python tools/bench_spmm_blocks.py --steps 10 > "$out/bench_spmm_blocks.jsonl" 2> "$out/bench_spmm_blocks.err"; echo "spmm blocks: exit $? [$(( $(date +%s) - t0 )) s]"
cut -c1-260 "$out/bench_spmm_blocks.jsonl"; tail -3 "$out/bench_spmm_blocks.err"
t0=$(date +%s)
timeout 1200 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench: exit $? [$(( $(date +%s) - t0 )) s]"
python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_n1.json') if l.startswith('{')][-1])
    lg=d.get('lightgcn') or {}
    print('bench: value %.3e e2e %.3e ms/step %.3f k1 %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['frac']))
    for k,v in lg.items():
        if k.startswith('batch'): print('   lightgcn', k, 'graph %.3f ms eager %.3f ms' % (v['ms_per_step'], v['ms_per_step_eager_launches']), v['cuda_graph'], v['graph_error'])
    print('   spmm', (lg.get('spmm') or {}).get('ms'), '| lightgcn error:', lg.get('error'))
    pc=d.get('parity_check') or {}
    if pc: print('   parity loss %.2e P %.3f Q %.3f' % (pc['loss_sum_neg_log_sigmoid']['rel_err'], pc['P']['rms_err_over_rms_update'], pc['Q']['rms_err_over_rms_update']))
    print('   clocks', d.get('clocks'))
except Exception as e:
    print('bench FAILED', e); print(open('$out/bench_n1.err').read()[-1500:])
PY
