#!/bin/bash
# round 2, call j: LJ-55 paired kernel, alternative thread mappings (PITA_LJ_CFG) at 1M and 4M configurations
cd "$GRAFT_REPO_ROOT"
for c in 0 3 4 1; do echo "cfg=$c"; PITA_LJ_CFG=$c timeout 300 python bench_lj.py --n 55 --batches 1048576,4194304 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('  B=%d ms=%.3f frac=%.3f finite=%s' % (d['batch'], d['ms'], d['frac_fp32_peak'], d['finite']))
    except Exception: print(l[:200])
"; done > gpurun_out/r2j_lj_cfgs.txt 2>&1
cat gpurun_out/r2j_lj_cfgs.txt
