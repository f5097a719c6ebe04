#!/usr/bin/env bash
# Last call of the round: exactly what the driver runs (GPU suite with -x, smoke, the default bench line, the reference arm).
set -u
out=gpurun_out/final_run
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; tail -20 "$out/build.log"; exit 1; }
timeout 1800 python -m pytest tests/ -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest -x -m gpu: exit $? -- $(tail -1 "$out/pytest_gpu.log")"
grep -E "^(FAILED|ERROR)" "$out/pytest_gpu.log" | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke: exit $? -- $(tail -1 "$out/smoke.log")"
t0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > "$out/bench.json" 2> "$out/bench.err"; echo "bench: exit $? wall $(( $(date +%s) - t0 )) s"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$out/bench_reference.json" 2> "$out/bench_reference.err"; echo "reference arm: exit $?"
timeout 600 python tools/bench_models.py --steps 5 > "$out/bench_models.jsonl" 2> "$out/bench_models.err"; echo "bench_models: exit $?"; cut -c1-200 "$out/bench_models.jsonl"
python - <<PY
import json
d=json.loads([l for l in open('$out/bench.json') if l.startswith('{')][-1])
r=json.loads([l for l in open('$out/bench_reference.json') if l.startswith('{')][-1])
print('value %.4e e2e %.4e ms %.3f | reference %.4e -> ratio e2e %.0f' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['value'], d['e2e']['value']/r['value']))
print('keys', sorted(d.keys()))
print('clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
