#!/usr/bin/env bash
# Last call of the round: exactly what the driver runs (GPU suite with -x, smoke, the default bench line, the reference arm).
set -u
out=gpurun_out/final_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
# the tensor-core K8 kernel first, on its own and with a short limit: a failure there must not take the suite with it
timeout 300 python -m pytest tests/test_gpu_topn.py -x -q -m gpu > "$out/pytest_topn.log" 2>&1; rc=$?; echo "pytest test_gpu_topn: exit $rc -- $(tail -1 "$out/pytest_topn.log")"
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error" "$out/pytest_topn.log" | head -8; export QREC_SKIP_TC=1; fi
timeout 1800 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest -x -m gpu: exit $? -- $(tail -1 "$out/pytest_gpu.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu.log" | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke: exit $? -- $(tail -1 "$out/smoke.log")"
t0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"; echo "bench: exit $? wall $(( $(date +%s) - t0 )) s"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "reference arm: exit $?"
timeout 600 python tools/bench_models.py --steps 5 > "$out/bench_models.jsonl" 2> "$out/bench_models.err"; echo "bench_models: exit $?"; cut -c1-200 "$out/bench_models.jsonl"
timeout 300 python - > "$out/k8_timing.log" 2>&1 <<'PY2'
import torch
from qrec_b200 import engine as E, synthetic
dev=torch.device('cuda',0)
data=synthetic.make_interactions(65536,100000,50,device=dev); P,Q=synthetic.init_tables(65536,100000,64,seed=1,device=dev)
users=torch.arange(65536,dtype=torch.int32,device=dev)
import os
for tc in ((False, True) if os.environ.get('QREC_SKIP_TC') != '1' else (False,)):
  for N in (10,100):
    E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N,tensor_cores=tc); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(3): E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N,tensor_cores=tc)
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/3
    print('K8 score_topn (%s) 65536 users x 100K items N=%d: %.2f ms = %.1f TFLOP/s (2*M*N*d), %.2f M users/s' % ('tcgen05 3xTF32' if tc else 'fp32 SIMT', N, ms, 65536*1e5*128/ms/1e9, 65536/ms/1e3))
PY2
cat "$out/k8_timing.log" | tail -5
timeout 600 ncu --clock-control none --set full --import-source on -k regex:score_topn -s 1 -c 1 -o "$out/topn_v2_full_r2" -f python tools/ncu_targets.py topn > "$out/ncu_topn.log" 2>&1
ncu -i "$out/topn_v2_full_r2.ncu-rep" --page raw --csv > "$out/topn_v2_full_r2_raw.csv" 2>/dev/null; echo "ncu topn: $(wc -c < "$out/topn_v2_full_r2_raw.csv") bytes"; rm -f "$out/topn_v2_full_r2.ncu-rep"
python - <<PY
import json
d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('$out/bench_reference.json') if l.startswith('{')][-1])
print('value %.4e e2e %.4e ms %.3f | reference %.4e -> ratio e2e %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['value'], d['e2e']['value']/r['value']))
print('keys', sorted(d.keys()))
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
