#!/usr/bin/env bash
# short single-GPU call: the suites touched since the last full run + the K8 timing of both kernels
set -u
out=gpurun_out/quick_check
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 300 python -m pytest tests/test_gpu_topn.py -x -q -m gpu > "$out/pytest_topn.log" 2>&1; rc=$?; echo "pytest test_gpu_topn: exit $rc -- $(tail -1 "$out/pytest_topn.log")"
if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR)|Error" "$out/pytest_topn.log" | head -8; export QREC_SKIP_TC=1; fi
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_models.py tests/test_gpu_dropin.py -q -m gpu > "$out/pytest_sel.log" 2>&1; echo "pytest graph+models+dropin: exit $? -- $(tail -1 "$out/pytest_sel.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_sel.log" | head
timeout 300 python - > "$out/k8_timing.log" 2>&1 <<'PY2'
import os, torch
from qrec_b200 import engine as E, synthetic
dev=torch.device('cuda',0)
data=synthetic.make_interactions(65536,100000,50,device=dev); P,Q=synthetic.init_tables(65536,100000,64,seed=1,device=dev)
users=torch.arange(65536,dtype=torch.int32,device=dev)
for tc in ((False, True) if os.environ.get('QREC_SKIP_TC') != '1' else (False,)):
  for N in (10,100):
    E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N,tensor_cores=tc); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(3): E.score_topn(P,Q,users,data['sorted_rowptr'],data['sorted_cols'],N,tensor_cores=tc)
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/3
    print('K8 score_topn (%s) 65536 users x 100K items N=%d: %.2f ms = %.1f TFLOP/s (2*M*N*d), %.2f M users/s' % ('tcgen05 3xTF32' if tc else 'fp32 SIMT', N, ms, 65536*1e5*128/ms/1e9, 65536/ms/1e3))
PY2
tail -5 "$out/k8_timing.log"
if [ "${QREC_SKIP_TC:-0}" != 1 ]; then
  QREC_TOPN_TC=1 timeout 600 ncu --clock-control none --set full --import-source on -k regex:score_topn_tc -s 1 -c 1 -o "$out/topn_tc_full_r2" -f python tools/ncu_targets.py topn > "$out/ncu_topn_tc.log" 2>&1
  ncu -i "$out/topn_tc_full_r2.ncu-rep" --page raw --csv > "$out/topn_tc_full_r2_raw.csv" 2>/dev/null; echo "ncu topn tc: $(wc -c < "$out/topn_tc_full_r2_raw.csv") bytes"
  ncu -i "$out/topn_tc_full_r2.ncu-rep" --page source --csv > "$out/topn_tc_full_r2_source.csv" 2>/dev/null; rm -f "$out/topn_tc_full_r2.ncu-rep"
fi
