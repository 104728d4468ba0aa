"""Print the per-rank table of a tools/shard_emulate.py JSON:  python tools/show_emul.py FILE"""
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:round(v,3) for k,v in d['unpartitioned_ms'].items()})
for W,rows in d['worlds'].items():
    print("world",W, "max step", round(max(r['step_ms'] for r in rows),3), "eff", round(d['unpartitioned_ms']['total']/int(W)/max(r['step_ms'] for r in rows),3))
    for r in rows:
        print({k.replace('upward_exchange','up').replace('near_field_join_l2p','join').replace('downward','down').replace('result_allreduce','res'):round(v,3) for k,v in r.items()})
