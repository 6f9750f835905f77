#!/bin/bash
# dev: S4 (wiki-103 shape) kernel matrix -> gpurun_out/s4_matrix.jsonl (one bench line per variant, tagged)
out=gpurun_out/s4_matrix.jsonl; : > $out
run() { tag=$1; shift; python bench.py --workload s4 --steps 40 --warmup 3 "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); d['tag']='$tag'; print(json.dumps(d))" >> $out; }
run e128_alias --window-refresh 0
run e128_alias_hot16 --window-refresh 0 --hot-rows 16
run e128_alias_hot48 --window-refresh 0 --hot-rows 48
run e128_alias_hot64_refresh --hot-rows 64
run e128_alias_refresh
run e48k3_alias_refresh --emb 48 --neg 3
run e48k3_alias_hot256_refresh --emb 48 --neg 3 --hot-rows 256
run e48k3_alias_hot1024_refresh --emb 48 --neg 3 --hot-rows 1024
run e64_alias_hot256_refresh --emb 64 --hot-rows 256
python - <<'PY'
import json
for l in open('gpurun_out/s4_matrix.jsonl'):
    d=json.loads(l)
    print(f"{d['tag']:24s} {d['value']/1e9:6.3f} G pairs/s  e2e {d['e2e']['value']/1e9:6.3f}  frac {d['roofline']['frac']:.3f}  loss {d['train_stats']['loss']:.4f}  hot {d['config'].get('hot_rows')} refresh {d['config'].get('window_refresh')}  {d['roofline']['kernel'][:60]}")
PY
