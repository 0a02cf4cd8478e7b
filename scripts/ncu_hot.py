"""Top stall-sampled SASS instructions (with context) of one kernel in an ncu report:
   python scripts/ncu_hot.py <prof.ncu-rep> <kernel regex> [top_n]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{rx}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
b = blocks[0]
idx = {h: i for i, h in enumerate(b['hdr'])}
data = b['data']
S = lambda r: int(r[idx['# Samples']] or 0)
tot = sum(S(r) for r in data)
print(b['name'][:80], '| total samples', tot, '| instructions', len(data))
stall_cols = [h for h in b['hdr'] if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall_cols}
print('stall reasons:', ', '.join(f'{h[6:]} {100 * v / max(tot, 1):.0f}%' for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for i in sorted(range(len(data)), key=lambda i: -S(data[i]))[:top_n]:
    print(f'---- {S(data[i])} samples ({100 * S(data[i]) / tot:.1f}%)')
    for k in range(max(0, i - 6), min(len(data), i + 2)):
        r = data[k]
        print('   ', str(S(r)).rjust(6), r[idx['Instructions Executed']].rjust(9), r[idx['Source']][:110])
