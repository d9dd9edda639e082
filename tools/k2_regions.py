"""Share of the executed warp instructions of the K2 filter kernel per source region.

    python tools/k2_regions.py <object file> <ncu --page source --csv output> <configs per launch>
"""
import sys, collections, csv, re
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_by_line as S
obj, src, ncfg = sys.argv[1], sys.argv[2], float(sys.argv[3])
table = S.line_table(obj,'fk_cc_filter_kernelILi6EjLb0ELb1ELi14E')
rows=[];hdr=None
for row in csv.reader(open(src, newline='')):
    if not row: continue
    if row[0]=='Address': hdr=row
    elif hdr and row[0].startswith('0x'): rows.append(row)
assert len(rows)==len(table), (len(rows), len(table))
iI=hdr.index('Instructions Executed'); iT=hdr.index('Thread Instructions Executed'); iS=hdr.index('Source')
# regions by source text markers in the current acro_filter.cuh
lines=open('/root/repo/acrobotics_b200/csrc/acro_filter.cuh').read().split('\n')
def find(sub, start=0):
    for i in range(start, len(lines)):
        if sub in lines[i]: return i+1
    raise KeyError(sub)
marks=[('sincos', find('void sincos_filter')), ('sat_filter', find('int sat_filter')), ('misc', find('template <typename MaskT>')),
       ('slotloop+joint', find('for (int slot = 0; slot <= last_slot')), ('boxpose', find('for (int b = rb.slot_begin[slot]')),
       ('grid lookup', find('// ---- candidates')), ('self cull', find('if (do_self) {', find('// ---- candidates'))),
       ('load_other', find('auto load_other')), ('round ctl', find('for (;;) {')), ('direct round', find('if (b2 == 0u ||')),
       ('redistribute', find('int incl = c;')), ('save pose', find('if (do_self && save_slot >= 0)')), ('end', find('retired = done;'))]
def region(ln):
    name='?'
    for n,l in marks:
        if ln>=l: name=n
    return name
tot=collections.Counter(); thr=collections.Counter(); ops=collections.defaultdict(collections.Counter)
for k,row in enumerate(rows):
    loc=table[k][2]; inst=int(row[iI]); t=int(row[iT])
    if loc and loc[0]=='acro_filter.cuh': name=region(loc[1])
    elif loc: name=loc[0]
    else: name='none'
    tot[name]+=inst; thr[name]+=t; ops[name][S.opcode(row[iS])]+=inst
T=sum(tot.values())
print(f"total warp inst/cfg {T/ncfg:.2f}  lanes {sum(thr.values())/T:.2f}")
for n,v in tot.most_common(25):
    top=' '.join(f"{o}:{100*c/v:.0f}" for o,c in ops[n].most_common(7))
    print(f"{n:22s} {100*v/T:6.2f}% {v/ncfg:6.2f}/cfg lanes {thr[n]/max(v,1):5.1f} | {top}")
