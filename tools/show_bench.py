"""Print the headline numbers, stage times and (N > 1) the per-rank table of a bench.py JSON line:  python tools/show_bench.py FILE"""
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d['value'],2), "ms", round(d['ms_per_step'],4), "e2e", round(d['e2e']['value'],2), "gpus", d['n_gpus'], "clocks", d.get('clocks'))
if 'partition' in d:
    for r in d['partition']['per_rank']: print({k.replace('_ms','').replace('kernel_','k_'):(round(v,3) if k!='rel_l2_vs_unpartitioned' else v) for k,v in r.items()})
else:
    print({k:round(v,3) for k,v in d['stages']['ms'].items()})
if 'fit' in d: print({k:v for k,v in d['fit'].items() if k!='workload'})
