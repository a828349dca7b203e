import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value %.1f MP/s  ms/step %.2f   e2e %.1f MP/s (%.1f ms)'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e'].get('ms_per_step',0)))
tot=0
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms_per_launch']*kv[1]['launches_per_step']):
    t=v['ms_per_launch']*v['launches_per_step']; tot+=t
    print('%-28s %9.3f ms x%-4.1f share %.4f  %s GB/s frac %s'%(k,v['ms_per_launch'],v['launches_per_step'],v['share_of_step'],v['algorithmic_gbs'],v['frac_of_peak']))
print('sum of kernels %.2f ms'%tot, '| launches', d['gpu_launches'], '| clocks', d['clocks'])
print(d.get('cpu_baseline')); print(d.get('parity'))
