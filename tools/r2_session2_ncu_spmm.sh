#!/usr/bin/env bash
# ncu --set full of the two bipartite halves of K2 inside a sharded LightGCN step (first layer: item side, then user side)
set -u
out=gpurun_out/r2s2d
mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo "build failed"; exit 1; }
timeout -k 5 100 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spmm_rowsplit -c 2 -o "$out/spmm_halves_full_r2s2" -f python tools/dist_lightgcn.py --skip-parity --profile-range --steps 1 > "$out/ncu.log" 2>&1; echo "ncu: exit $?"
ncu -i "$out/spmm_halves_full_r2s2.ncu-rep" --page raw --csv > "$out/spmm_halves_full_r2s2_raw.csv" 2>/dev/null; echo "raw csv: $(wc -c < "$out/spmm_halves_full_r2s2_raw.csv") bytes"
rm -f "$out/spmm_halves_full_r2s2.ncu-rep"
python tools/ncu_summary.py "$out/spmm_halves_full_r2s2_raw.csv" 2>&1 | head -30
