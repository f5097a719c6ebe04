#!/usr/bin/env bash
# round 2, second session, N-GPU call (N from env, default 2): the sharded LightGCN step replayed from a CUDA graph with the
# NCCL all-reduces captured -- parity against the single-GPU class, then the bench line (graph vs eager step times)
set -u
N=${N:-2}
out=gpurun_out/r2s2_n$N
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29911 tools/dist_lightgcn.py --steps 10 --graph > "$out/dist_lightgcn_graph.log" 2>&1; echo "dist_lightgcn --graph: exit $? -- $(grep -h '^{' "$out/dist_lightgcn_graph.log" | cut -c1-500)"; grep -E "AssertionError|Error" "$out/dist_lightgcn_graph.log" | head -5
timeout 900 $TR --master-port 29912 bench.py --gpus $N --steps 20 --warmup 5 --no-neumf --no-extras --no-roofs ${BENCH_FLAGS:---no-parity-multi} > "$out/bench.json" 2> "$out/bench.err"; echo "bench: exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
    lg=d.get('lightgcn') or {}
    print('bench N=$N: value %.3e e2e %.3e ms/step %.3f k1 %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms']))
    for k,v in lg.items():
        if k.startswith('batch'): print('   lightgcn', k, 'chosen %.3f ms | graph' % v['ms_per_step'], v['ms_per_step_graph_replay'], '| eager %.3f ms' % v['ms_per_step_eager_launches'], v['cuda_graph'], v['graph_error'])
    print('   lightgcn error:', lg.get('error'))
except Exception as e:
    print('bench FAILED', e); print(open('$out/bench.err').read()[-1500:])
PY
