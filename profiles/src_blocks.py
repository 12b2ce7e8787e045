"""Source page of an ncu report (ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv) grouped into blocks of SASS rows
with equal execution counts: warp instructions, share, active lanes, stall samples.  usage: python profiles/src_blocks.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; iS = h.index('Source'); iE = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iSm = h.index('# Samples')
data = [(r[iS].strip(), int(r[iE]), int(r[iT]), int(r[iSm])) for r in rows[2:] if len(r) > iT]
tot = sum(d[1] for d in data); tots = sum(d[3] for d in data)
print("# %s" % rows[0][1][:90])
print("# total warp instructions %.0f M, threads active per instruction %.1f, %d SASS rows" % (tot / 1e6, sum(d[2] for d in data) / tot, len(data)))
g = []
for idx, d in enumerate(data):
    if g and abs(d[1] - g[-1]['e']) <= 0.03 * max(d[1], g[-1]['e']):
        g[-1]['n'] += 1; g[-1]['w'] += d[1]; g[-1]['t'] += d[2]; g[-1]['s'] += d[3]; g[-1]['last'] = idx
    else:
        g.append({'first': idx, 'last': idx, 'n': 1, 'e': d[1], 'w': d[1], 't': d[2], 's': d[3]})
for x in g:
    if x['w'] / tot > 0.004:
        print("rows %3d-%3d n=%3d executions %7.2fM  warp instructions %7.1fM (%4.1f%%)  threads active %4.1f  stall samples %4.1f%%  first: %s"
              % (x['first'], x['last'], x['n'], x['e'] / 1e6, x['w'] / 1e6, 100 * x['w'] / tot, x['t'] / max(1, x['w']), 100 * x['s'] / tots, data[x['first']][0][:44]))
