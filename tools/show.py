"""print the headline fields of bench.py JSON lines: python tools/show.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        s = d.get("sustained") or {}
        print(f"{f}: N={d['n_gpus']} value {d['value']:.1f} ({d['ms_per_step']:.2f} ms/step) e2e {d['e2e']['value']:.1f} | sustained {s.get('value')} e2e {s.get('e2e_value')} | "
              f"probe {d.get('parity_probe', {}) and d['parity_probe'].get('ok')} | {d['config'].get('stage_ms_per_step_serial')}")
    except Exception as e:
        print(f, "unreadable:", e)
