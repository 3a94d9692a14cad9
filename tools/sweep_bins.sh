#!/usr/bin/env bash
# sweep of the average bin size (HSK_TARGET_BIN) on both bench workloads
for t in 2048 3072 4096 6144 8192 12288; do
  for w in c2_150Mbp_10kbp c2_150Mbp_150bp; do
    HSK_TARGET_BIN=$t python bench.py --workload $w --no-cpu-baseline --steps 8 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']
print('target $t', '$w', 'G/s %.2f'%(d['value']/1e9), 'e2e %.2f'%(d['e2e']['value']/1e9), 'ms', round(d['ms_per_step'],2), 'extract %.2f bins %.2f'%(s['ms_extract'], s['ms_bins']), 'ovf', d['config']['overflow_bins'])"
  done
done
