import json, sys
for w in ('c2tc', 'c3', 'c4'):
    try:
        d = json.loads(open('gpurun_out/chk_%s.json' % w).read().strip().splitlines()[-1])
        print(w, '%.3e' % d['value'], '%.1f us' % (d['ms_per_step'] * 1e3), d.get('step_kernels_us'))
    except Exception as e:
        print(w, 'err', e)
        try: print(open('gpurun_out/chk_%s.err' % w).read()[-1500:])
        except Exception: pass
