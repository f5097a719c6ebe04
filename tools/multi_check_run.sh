#!/usr/bin/env bash
# quick multi-GPU re-check (N from env): data-parallel NeuMF, sharded LightGCN / SimGCL with the row-restricted layers, bench line
set -u
N=${N:-2}
out=gpurun_out/multi_check_n$N
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29900
port=$((port+1)); timeout 600 $TR --master-port $port tools/dist_neumf.py --steps 10 > "$out/dist_neumf.log" 2>&1; echo "dist_neumf: exit $? -- $(grep -h '^{' "$out/dist_neumf.log" | cut -c1-900)"; grep -E "AssertionError|Error" "$out/dist_neumf.log" | head -3
[ "${SKIP_DIST_LGCN:-0}" = 1 ] || { port=$((port+1)); timeout 600 $TR --master-port $port tools/dist_lightgcn.py --steps 10 > "$out/dist_lightgcn.log" 2>&1; echo "dist_lightgcn: exit $? -- $(grep -h '^{' "$out/dist_lightgcn.log" | cut -c1-400)"; grep -E "AssertionError|Error" "$out/dist_lightgcn.log" | head -3; }
[ "${SKIP_DIST_LGCN:-0}" = 1 ] || { port=$((port+1)); timeout 600 $TR --master-port $port tools/dist_lightgcn.py --steps 10 --scheme cols > "$out/dist_lightgcn_cols.log" 2>&1; echo "dist_lightgcn cols: exit $? -- $(grep -h '^{' "$out/dist_lightgcn_cols.log" | cut -c1-400)"; grep -E "AssertionError|Error" "$out/dist_lightgcn_cols.log" | head -3; }
[ "${SKIP_DIST_SIMGCL:-0}" = 1 ] || { port=$((port+1)); timeout 900 $TR --master-port $port tools/dist_simgcl.py --steps 5 > "$out/dist_simgcl.log" 2>&1; echo "dist_simgcl: exit $? -- $(grep -h '^{' "$out/dist_simgcl.log" | cut -c1-600)"; grep -E "AssertionError|Error" "$out/dist_simgcl.log" | head -3; }
[ "${SKIP_BENCH:-0}" = 1 ] && exit 0
for w in ${WAVES:-1 2}; do
port=$((port+1)); timeout 600 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --q-syncs $w > "$out/bench_w$w.json" 2> "$out/bench_w$w.err"
python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_w$w.json') if l.startswith('{')][-1])
    pc=d.get('parity_check') or {}; lg=d.get('lightgcn') or {}
    print('bench w$w: value %.3e e2e %.3e ms/step %.3f k1 %.3f | loss %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['launch_ms'], d['config']['epoch_loss']))
    if pc: print('   parity loss %.2e P %.3f Q %.3f' % (pc['loss_sum_neg_log_sigmoid']['rel_err'], pc['P']['rms_err_over_rms_update'], pc['Q']['rms_err_over_rms_update']))
    if lg: print('   lightgcn', {k:round(v['ms_per_step'],3) for k,v in lg.items() if k.startswith('batch')})
except Exception as e:
    print('bench w$w FAILED', e); print(open('$out/bench_w$w.err').read()[-1200:])
PY
done

# LightGCN section alone with the feature-parallel scheme
port=$((port+1)); QREC_LGCN_SCHEME=cols timeout 600 $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-parity --no-roofs > "$out/bench_cols.json" 2> "$out/bench_cols.err"
python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_cols.json') if l.startswith('{')][-1])
    lg=d.get('lightgcn') or {}
    print('bench (lightgcn scheme %s): value %.3e | lightgcn' % (lg.get('scheme'), d['value']), {k:round(v['ms_per_step'],3) for k,v in lg.items() if k.startswith('batch')})
except Exception as e:
    print('bench cols FAILED', e); print(open('$out/bench_cols.err').read()[-1200:])
PY
