import json, sys
for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.3e ms/step %.3f e2e %.3e whole-step frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step']['frac']))
    for k, v in d['roofline']['kernels'].items():
        print('   %-40s %8.1f us %7.0f GB/s frac %.3f' % (k, v['us'], v['GBps'], v['frac']))
