#!/usr/bin/env bash
# First job of the next round (one gpurun call, ~10 GPU-minutes): run everything that was written after
# round 1's GPU budget was spent and is still gated by QREC_TEST_UNVALIDATED, each suite in its own
# process (a kernel fault poisons only its own CUDA context), then the side-by-side timings.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/first_hardware_run.sh'
# Results land in gpurun_out/first_run/.
set -u
export QREC_TEST_UNVALIDATED=1
out=gpurun_out/first_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
for t in test_gpu_table_sync test_gpu_topn test_gpu_adjacency test_gpu_k1_tma test_gpu_k1_sig test_gpu_rating test_gpu_lightgcn_blocked test_gpu_spmm_variants test_gpu_tcgemm_v2 test_gpu_parity_config2; do
  timeout 600 python -m pytest "tests/$t.py" -m gpu -q -s > "$out/$t.log" 2>&1
  echo "$t: exit $? -- $(tail -1 "$out/$t.log")"
done
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_all.py > "$out/sanitizer_r2_memcheck.log" 2>&1; echo "memcheck: exit $? -- $(grep -c "ERROR SUMMARY" "$out/sanitizer_r2_memcheck.log") $(grep "ERROR SUMMARY" "$out/sanitizer_r2_memcheck.log" | tail -1)"
timeout 600 python tools/bench_k1.py      > "$out/bench_k1.jsonl"      2> "$out/bench_k1.err";      echo "bench_k1: exit $?"
timeout 600 python tools/bench_rating.py  > "$out/bench_rating.jsonl"  2> "$out/bench_rating.err";  echo "bench_rating: exit $?"
timeout 600 python tools/bench_graph.py --spmm-only > "$out/bench_spmm.jsonl" 2> "$out/bench_spmm.err"; echo "bench_spmm: exit $?"
timeout 600 python tools/bench_tcgemm.py  > "$out/bench_tcgemm.jsonl"  2> "$out/bench_tcgemm.err";  echo "bench_tcgemm: exit $?"
for b in 1 3; do
  QREC_LGCN_ITEM_BLOCKS=$b timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > "$out/bench_lgcn_blocks$b.json" 2> "$out/bench_lgcn_blocks$b.err"
  echo "lightgcn item blocks $b: $(python -c "import json,sys; d=json.load(open('$out/bench_lgcn_blocks$b.json')); print(d.get('lightgcn'))" 2>/dev/null | cut -c1-300)"
done
grep -h "k1_variant\|k9\|tc_gemm\|spmm" "$out"/*.jsonl | cut -c1-220
timeout 900 python bench.py --steps 10 --warmup 3 --no-lightgcn > "$out/bench_parity.json" 2> "$out/bench_parity.err"; echo "bench parity: exit $?"
python -c "import json; d=json.load(open('$out/bench_parity.json')); print(json.dumps(d.get('parity_check'), indent=1)[:3000]); print(d['value'], d['e2e']['value'])"
