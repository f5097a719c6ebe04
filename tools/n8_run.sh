#!/usr/bin/env bash
# Round 2, the one 8-GPU call: BPR scaling with the old and the new item-table exchange, LightGCN at N=8, config 5.
set -u
N=8
out=gpurun_out/multi_n8
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29800
run_bench() {
  port=$((port+1))
  timeout 600 $TR --master-port $port bench.py --gpus $N --steps 20 --warmup 5 $2 > "$out/bench_$1.json" 2> "$out/bench_$1.err"
  rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_$1.json') if l.startswith('{')][-1])
    pc=d.get('parity_check') or {}
    lg=d.get('lightgcn') or {}
    print('$1 rc=$rc value %.3e e2e %.3e ms/step %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'][:170]))
    if pc: print('   parity loss %.2e P %.3f Q %.3f replicas_equal %s' % (pc['loss_sum_neg_log_sigmoid']['rel_err'], pc['P']['rms_err_over_rms_update'], pc['Q']['rms_err_over_rms_update'], pc.get('item_table_replicas_bit_identical_after_drain')))
    if lg: print('   lightgcn', {k:(round(v['ms_per_step'],3), v['loss']) for k,v in lg.items() if k.startswith('batch')})
except Exception as e:
    print('$1 rc=$rc FAILED to parse:', e); print(open('$out/bench_$1.err').read()[-1500:])
PY
}
run_bench blocking "--qsync blocking --no-parity-multi --no-lightgcn-multi"
run_bench p2p "--qsync p2p --parity-multi --lightgcn-multi"
run_bench nccl "--qsync nccl --no-parity-multi --no-lightgcn-multi"
run_bench p2p_w4 "--qsync p2p --q-syncs 4 --no-parity-multi --no-lightgcn-multi"
run_bench p2p_w1 "--qsync p2p --q-syncs 1 --no-parity-multi --no-lightgcn-multi"
QREC_PEER_ALLREDUCE=1 run_bench lgcn_peer "--qsync p2p --no-parity --lightgcn-multi"
port=$((port+1)); timeout 900 $TR --master-port $port tools/dist_simgcl.py --steps 5 --skip-parity > "$out/dist_simgcl.log" 2>&1; echo "dist_simgcl (config 5): exit $? -- $(grep -h '^{' "$out/dist_simgcl.log" | cut -c1-900)"
port=$((port+1)); QREC_PEER_ALLREDUCE=1 timeout 900 $TR --master-port $port tools/dist_simgcl.py --steps 5 --skip-parity > "$out/dist_simgcl_peer.log" 2>&1; echo "dist_simgcl peer (config 5): exit $? -- $(grep -h '^{' "$out/dist_simgcl_peer.log" | cut -c1-400)"
port=$((port+1)); timeout 600 $TR --master-port $port tools/dist_bpr_sharded.py --steps 5 --minibatch 1048576 --batch 4194304 > "$out/dist_bpr_sharded.log" 2>&1; echo "dist_bpr_sharded: exit $? -- $(grep -h '^{' "$out/dist_bpr_sharded.log" | cut -c1-400)"
port=$((port+1)); timeout 600 $TR --master-port $port tools/dist_neumf.py --steps 10 > "$out/dist_neumf.log" 2>&1; echo "dist_neumf: exit $? -- $(grep -h '^{' "$out/dist_neumf.log" | cut -c1-700)"; tail -3 "$out/dist_neumf.log" | grep -v '^{' | tail -2
