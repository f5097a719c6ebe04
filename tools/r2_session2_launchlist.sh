#!/usr/bin/env bash
# launch list (ncu gpu__time_duration, cold-cache, serialised) of ONE sharded LightGCN minibatch step at world 1
set -u
out=gpurun_out/r2s2c
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
timeout -k 5 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file "$out/launches_lgcn_step.csv" python tools/dist_lightgcn.py --skip-parity --profile-range --steps 1 > "$out/ncu_lgcn.log" 2>&1; echo "ncu lgcn: exit $? $(wc -l < "$out/launches_lgcn_step.csv") lines"
python - <<PY
import csv, collections
try:
    rows=[r for r in csv.reader(l for l in open('$out/launches_lgcn_step.csv') if l.startswith('"'))]
    hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
    agg=collections.OrderedDict()
    for r in rows[1:]:
        v=float(r[vi].replace(',','')); v = v/1e3 if r[ui] in ('ns','nsecond') else v
        k=r[ki][:70]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    print('launch list of 1 step: %d launches, %.1f us' % (sum(a[0] for a in agg.values()), tot))
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:24]: print('  %-70s x%-3d %9.1f us  %.1f%%' % (k,a[0],a[1],100*a[1]/tot))
except Exception as e:
    print('launch list parse failed', e)
PY
