#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python tools/sweep.py --workloads products --windows -1 --F 8,16,24,32 --us 2,4,18,20 > gpurun_out/sweep_narrow.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json,collections
rows=[json.loads(l) for l in open('gpurun_out/sweep_narrow.log') if l.startswith('{') and '"G"' in l and 'BEST' not in l]
grp=collections.defaultdict(list)
for r in rows: grp[(r['F'],r['weighted'])].append(r)
for k,v in sorted(grp.items()):
    v.sort(key=lambda r:r['ms'])
    print(k, ' | '.join(f"G{r['G']}V{r['V']}U{r['U']}:{r['ms']:.3f}" for r in v[:6]))
for l in open('gpurun_out/sweep_narrow.log'):
    if 'BEST' in l:
        d=json.loads(l); print('default', d['BEST']['F'], d['BEST']['weighted'], d['default_ms'])
PY
