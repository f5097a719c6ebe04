#!/usr/bin/env bash
# Round 2, second GPU call: re-run the suites that failed in the first call, the single-sweep K1, the whole bench line.
set -u
out=gpurun_out/second_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
export QREC_TEST_UNVALIDATED=1
for t in test_gpu_k1_sig test_gpu_rating test_gpu_tcgemm_v2 test_gpu_parity_config2 test_gpu_bpr test_gpu_k1_tma; do
  timeout 900 python -m pytest "tests/$t.py" -m gpu -q -s > "$out/$t.log" 2>&1
  echo "$t: exit $? -- $(tail -1 "$out/$t.log")"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke: exit $? -- $(tail -1 "$out/smoke.log")"
timeout 600 python tools/bench_k1.py > "$out/bench_k1.jsonl" 2> "$out/bench_k1.err"; echo "bench_k1: exit $?"; grep k1_variant "$out/bench_k1.jsonl" | cut -c1-200
timeout 1200 python bench.py --steps 20 --warmup 5 > "$out/bench_full.json" 2> "$out/bench_full.err"; echo "bench full: exit $?"
python - <<PY
import json
d=json.load(open('$out/bench_full.json'))
pc=d.get('parity_check',{})
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])
print('parity loss',pc.get('loss_sum_neg_log_sigmoid'),'P',pc.get('P'),'Q',pc.get('Q'))
print('roofline', json.dumps({k:v for k,v in d['roofline'].items() if k not in ('row_op_peak',)})[:900])
print('row_op', json.dumps(d['roofline'].get('row_op_peak'))[:1500])
for k in ('lightgcn','neumf','config1_filmtrust','zipf_contended','cpu_baseline','shuffled_order'):
    print(k, json.dumps(d.get(k))[:1200])
PY
QREC_K1_UM_CAP=8 timeout 600 python bench.py --steps 10 --warmup 3 --no-lightgcn --no-neumf --no-extras --no-roofs > "$out/bench_cap8.json" 2> "$out/bench_cap8.err"
python -c "import json; d=json.load(open('$out/bench_cap8.json')); pc=d['parity_check']; print('cap8 value',d['value'],'P',pc['P'],'Q',pc['Q'])"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_all.py > "$out/sanitizer_r2_memcheck.log" 2>&1; echo "memcheck: exit $? -- $(grep "ERROR SUMMARY" "$out/sanitizer_r2_memcheck.log" | tail -1)"
