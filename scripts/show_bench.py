import sys,json
for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'Gs/s %.1f ms %.4f e2e %.2f (%.2f ms) pipe_frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['pipeline']['frac_of_peak']), [(k['kernel'],round(k['ms'],4), round(k['algo_bytes']/k['ms']/1e6,0)) for k in d['pipeline']['kernels']], d.get('clocks'))
