#!/usr/bin/env bash
# N-GPU re-check of the process exit after a captured NCCL graph (parallel.finish_process): the bench line, strict timeouts
set -u
N=${N:-2}
out=gpurun_out/r2s2_n${N}b
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
t0=$(date +%s)
timeout -k 5 240 $TR --master-port 29921 bench.py --gpus $N --steps 20 --warmup 5 --no-neumf --no-extras --no-roofs --no-parity-multi > "$out/bench.json" 2> "$out/bench.err"; echo "bench: exit $? after $(( $(date +%s) - t0 )) s"
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
