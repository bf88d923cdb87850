import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    c = d['cycles_per_block']
    print(f.split('/')[-1], 'ms/step', round(d['ms_per_step'], 3), 'chain M/s', round(d['chain_snp_updates_per_s'] / 1e6, 2), 'ident', d['ranks_bit_identical'],
          'rows/w', d['config']['rows_per_worker_max'], 'workers', d['config']['workers'], 'kms', {k: round(v, 2) for k, v in d['kernel_ms_per_step'].items()})
    print('    gather1', round(c['gather_first_chunk']), 'serial', round(c['serial_pass']), 'wwait', round(c['worker_wait']), 'wdots', round(c['worker_dots']),
          'wred', round(c['worker_reduce']), 'chg', round(d['state_changing_marker_fraction'], 3), 'e2e', (d.get('e2e') or {}).get('value'))
