"""One-line digest of a bench.py JSON line.  Usage: python tools/bench_brief.py bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = d.get("roofline", {}).get("stage_ms", {})
print(f"{d['config'].get('workload')} n_gpus={d['n_gpus']} value={d['value']/1e9:.2f} G/s e2e={d['e2e']['value']/1e9:.2f} G/s "
      f"ms/step={d['ms_per_step']:.3f} extract={st.get('ms_extract', 0):.3f} exch={st.get('ms_exchange', 0):.3f} "
      f"bins={st.get('ms_bins', 0):.3f} ovf={d['config'].get('overflow_bins')} clocks={d.get('clocks', {}).get('sm_mhz')}")
