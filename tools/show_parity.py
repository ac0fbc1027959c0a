import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    p=json.loads(line)['parity']
    print(sys.argv[1], {k:(round(v,7) if isinstance(v,float) else v) for k,v in p.items() if k in ('grad_rel_l2','max_grad_rel','worst_gradient','ok','ranks')})
    print('   ', ' '.join('%s=%.1e' % (k.replace('/weights','W').replace('/bias','b'), v['cuda_vs_f64']) for k,v in p['gradients'].items()))
