#!/usr/bin/env bash
# Round 2 multi-GPU validation (gpurun --gpus N): the overlapped item-table exchange in all its backends with the
# full-epoch parity check, LightGCN with the peer-memory all-reduce, SimGCL over the sharded user table, K7.
#   N=${N:-2} bash tools/multi_gpu_run.sh
set -u
N=${N:-2}
out=gpurun_out/multi_n$N
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
run_bench() {   # name, extra args
  port=$((port+1))
  timeout 900 $TR --master-port $port bench.py --gpus $N --steps 10 --warmup 3 $2 > "$out/bench_$1.json" 2> "$out/bench_$1.err"
  rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_$1.json') if l.startswith('{')][-1])
    pc=d.get('parity_check') or {}
    lg=d.get('lightgcn') or {}
    print('$1 rc=$rc value %.3e e2e %.3e ms/step %.3f | %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['config']['parallelism'][:150]))
    if pc: print('   parity loss %.2e P %.3f Q %.3f replicas_equal %s' % (pc['loss_sum_neg_log_sigmoid']['rel_err'], pc['P']['rms_err_over_rms_update'], pc['Q']['rms_err_over_rms_update'], pc.get('item_table_replicas_bit_identical_after_drain')))
    if lg: print('   lightgcn', {k:(round(v['ms_per_step'],3), v['loss']) for k,v in lg.items() if k.startswith('batch')}, lg.get('impl','')[-80:])
except Exception as e:
    print('$1 rc=$rc FAILED to parse:', e); print(open('$out/bench_$1.err').read()[-1500:])
PY
}
run_bench blocking "--qsync blocking --parity-multi --no-lightgcn-multi"
run_bench nccl "--qsync nccl --parity-multi --no-lightgcn-multi"
run_bench p2p "--qsync p2p --parity-multi --no-lightgcn-multi"
run_bench p2p_w4 "--qsync p2p --q-syncs 4 --parity-multi --no-lightgcn-multi"
run_bench p2p_w1 "--qsync p2p --q-syncs 1 --parity-multi --no-lightgcn-multi"
run_bench lgcn_peer "--qsync p2p --no-parity --lightgcn-multi"
QREC_PEER_ALLREDUCE=0 run_bench lgcn_nccl "--qsync p2p --no-parity --lightgcn-multi"
port=$((port+1)); timeout 900 $TR --master-port $port tools/dist_lightgcn.py --steps 5 > "$out/dist_lightgcn.log" 2>&1; echo "dist_lightgcn: exit $? -- $(grep -h '^{' "$out/dist_lightgcn.log" | cut -c1-300)"
port=$((port+1)); timeout 1200 $TR --master-port $port tools/dist_simgcl.py --steps 5 > "$out/dist_simgcl.log" 2>&1; echo "dist_simgcl: exit $? -- $(grep -h '^{' "$out/dist_simgcl.log" | cut -c1-700)"; tail -5 "$out/dist_simgcl.log" | grep -v '^{' | tail -3
port=$((port+1)); timeout 900 $TR --master-port $port tools/dist_bpr_sharded.py --steps 5 > "$out/dist_bpr_sharded.log" 2>&1; echo "dist_bpr_sharded: exit $? -- $(grep -h '^{' "$out/dist_bpr_sharded.log" | cut -c1-400)"; tail -4 "$out/dist_bpr_sharded.log" | grep -v '^{' | tail -3
