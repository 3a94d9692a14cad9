"""One-line digest of a bench.py JSON line.  Usage: python tools/bench_brief.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = d.get("roofline", {}).get("stage_ms", {})
g = lambda x: f"{x['value']/1e9:.2f}" if x and x.get("value") else "-"
print(f"{d['config'].get('workload')} k={d['config'].get('k')} ext={d['config'].get('ext')} n_gpus={d['n_gpus']} value={d['value']/1e9:.2f} G/s "
      f"e2e(api)={g(d.get('e2e'))} e2e(pinned soa)={g(d.get('e2e_pinned_soa'))} ceiling={g(d.get('e2e_ceiling'))} G/s "
      f"ms/step={d['ms_per_step']:.3f} extract={st.get('ms_extract', 0):.3f} exch={st.get('ms_exchange', 0):.3f} "
      f"bins={st.get('ms_bins', 0):.3f} ovf={d.get('details', {}).get('overflow_bins')} parity={d.get('parity_check', {}).get('status')} "
      f"clocks={(d.get('clocks') or {}).get('sm_mhz')} frac={d.get('roofline', {}).get('frac', 0):.4f}")
