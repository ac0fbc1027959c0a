import json
import sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    try:
        d = json.loads(line)
    except Exception:
        start = line.find('{"metric"')
        if start < 0:
            continue
        try:
            d = json.loads(line[start:])
        except Exception:
            continue
    p = d.get('parity') or {}
    print('N={} {} {:.0f} samples/s {:.3f} ms/step  e2e {:.0f}  parity ok={} grad_l2={} (f32 oracle {})'.format(
        d.get('n_gpus'), d.get('scaling'), d.get('value', 0), d.get('ms_per_step', 0), d.get('e2e', {}).get('value', 0),
        p.get('ok'), p.get('grad_rel_l2'), p.get('grad_rel_l2_of_cpu_float32_oracle')))
