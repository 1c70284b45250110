"""Print the interesting parts of a bench.py JSON line (multi-GPU or single)."""
import json, sys
d = json.load(open(sys.argv[1]))
print('N', d['n_gpus'], 'value %.3e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'first %.4f' % d['first_call']['ms_per_step'], {k: round(v, 4) for k, v in d['kernel_ms'].items()})
print('e2e ms %.2f public ctor ms %.2f' % (d['e2e']['ms_per_step'], d['e2e']['public_constructor_wall_ms']), 'fp64', {k: round(v, 3) for k, v in d['roofline_fp64'].items() if isinstance(v, float)})
p = d['parity']
print('parity total', p['mismatches'], {k: (v.get('mismatches') if isinstance(v, dict) else v) for k, v in p.items()})
for k, v in p.items():
    if isinstance(v, dict):
        print('   ', k, {kk: (vv.get('mismatches') if isinstance(vv, dict) else None) for kk, vv in v.items() if isinstance(vv, dict)})
if 'partition_sweep' in d:
    print('sweep', {k: round(v['ms_per_step'], 4) for k, v in d['partition_sweep'].items()})
z = d['zipper']
if 'fold_exchange' in z:
    fx = z['fold_exchange']
    if 'error' in fx: print('fold error', fx['error'])
    for k, v in fx.items():
        if isinstance(v, dict) and k != 'parity':
            print(' ', k, {kk: (round(vv['ms_per_exchange'] * 1e3, 1), round(vv['gbs_per_direction_per_gpu'])) for kk, vv in v.items() if isinstance(vv, dict)},
                  {kk: round(vv, 1) for kk, vv in v.items() if isinstance(vv, float)}, v.get('error', ''))
    df = z['distributed_fill_halo_regions']
    print(' dist fill', df.get('partition'), {k: round(v['ms_per_fill'] * 1e3, 1) for k, v in df.items() if isinstance(v, dict) and 'ms_per_fill' in v}, df.get('error', ''))
else:
    print({k: (round(v['ms'] * 1e3, 1), round(v['gbs'])) for k, v in z.items() if isinstance(v, dict) and 'ms' in v})
    print('roofline_zipper', json.dumps(d['roofline_zipper'])[:600])
