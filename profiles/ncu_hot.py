"""Hot instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME` output:
samples, executed count and the three main stall reasons per SASS instruction (top N by samples, program order)."""
import csv
import sys


def main(path, n=40):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    data = [r for r in rows[1:] if r[0].startswith('0x')]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print('total samples', tot, 'instructions', len(data))
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {}
    for r in data:
        for h in stalls:
            agg[h[6:]] = agg.get(h[6:], 0) + int(r[ix[h]])
    print('stall totals', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    top = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:n]
    for i in sorted(top):
        r = data[i]
        br = sorted([(int(r[ix[h]]), h[6:]) for h in stalls if int(r[ix[h]]) > 0], reverse=True)[:3]
        print(i, r[ix['Source']].strip()[:48].ljust(48), r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(10), br)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
