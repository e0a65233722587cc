"""Print the key numbers of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value'] / 1e6:.1f} M/s  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value'] / 1e6:.1f} M/s  host {d.get('host_enqueue_ms_per_step', 0):.4f} ms  "
      f"fwd {d['roofline']['kernel_ms']}  frac {d['roofline']['frac']:.3f}  clocks {d['clocks']}")
