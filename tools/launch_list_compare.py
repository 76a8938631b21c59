"""Side-by-side per-launch table of ncu launch lists (CSV files of ONE bench step each): time, SM cycles, tensor-pipe %.\n    python tools/launch_list_compare.py profiles/r02c_pair_launches_fp16x3_one_cta.csv profiles/r02c_pair_launches_fp16x3_all_layers_as_pairs.csv"""
import csv,collections,sys
def load(f):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    d=collections.OrderedDict()
    for r in rows:
        d.setdefault(int(r[0]),{'k':r[4],'grid':r[8]})[r[12]]=float(r[14].replace(',',''))
    ids=sorted(d); st=[i for i in ids if 'stem_s2d' in d[i]['k']][0]
    return [d[i] for i in ids if i>=st]+[d[i] for i in ids if i<st]
names=sys.argv[1:]
L=[load(f'{n}') for n in names]
T='gpu__time_duration.sum'; C='sm__cycles_elapsed.max'; P='sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
tot=[0]*len(L)
for n in range(len(L[0])):
    row=[l[n] for l in L]
    for i,r in enumerate(row): tot[i]+=r[T]
    if 'conv_tc' not in row[0]['k']: continue
    print(f"{n:2d} "+" | ".join(f"{r[T]/1e3:7.1f}us {r[C]/1e3:6.0f}kc {r[P]:5.1f}% {r['k'][19:32]}" for r in row))
print('totals ms', [round(t/1e6,3) for t in tot])
