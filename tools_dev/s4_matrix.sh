#!/bin/bash
# dev: S4 (wiki-103 shape) kernel matrix -> gpurun_out/s4_matrix.jsonl (one bench line per variant, tagged).
# (The hot-row combining variants of profiles/r02_s4_matrix.md section 2 were removed from the library together with their --hot-rows flag.)
out=gpurun_out/s4_matrix.jsonl; : > $out
run() { tag=$1; shift; python bench.py --workload s4 --steps 40 --warmup 3 "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); d['tag']='$tag'; print(json.dumps(d))" >> $out; }
run e128_alias --window-refresh 0
run e128_alias_refresh
run e48k3_alias_refresh --emb 48 --neg 3
python - <<'PY'
import json
for l in open('gpurun_out/s4_matrix.jsonl'):
    d=json.loads(l)
    print(f"{d['tag']:24s} {d['value']/1e9:6.3f} G pairs/s  e2e {d['e2e']['value']/1e9:6.3f}  frac {d['roofline']['frac']:.3f}  loss {d['train_stats']['loss']:.4f}  refresh {d['config'].get('window_refresh')}  {d['roofline']['kernel'][:60]}")
PY
