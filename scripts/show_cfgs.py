import json
d=json.load(open('gpurun_out/time_configs.json'))
for k,r in d.items():
    print(k, r['out_shape'], 'ours %.3f ms (%.1f Gs/s)'%(r['ours_ms_best'], r['ours_gsamples']), 'ref_gpu', r.get('ref_gpu_ms_best'), r.get('ref_gpu_error','')[:60], 'x%.1f'%r.get('speedup_vs_ref_gpu',0), 'fft', r.get('fft_size'), 'ws %.1f GiB'%r.get('workspace_gib',0))
    for kk in r.get('kernels',[]): print('     ', kk['kernel'], '%.4f ms'%kk['ms'], '%.0f GB/s'%(kk['gbs'] or 0))
