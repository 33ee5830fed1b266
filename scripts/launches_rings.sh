#!/bin/bash
# per-kernel device times of the ring/signature phase (dense variant) at C2 and C3
for a in "20000 3" "100000 4"; do
  set -- $a
  python scripts/time_c2.py $1 $2 > gpurun_out/plain_rings_$1.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum --clock-control none -k regex:'ball|ring_cdf|bfs_ring' -s 20 -c 20 --csv \
      --log-file gpurun_out/launches_rings_$1.csv python scripts/time_c2.py $1 $2 > /dev/null 2>&1
  python - <<PY
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_rings_$1.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ix={h:i for i,h in enumerate(hdr)}
rec=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)!=len(hdr): continue
    rec.setdefault((r[ix['ID']], r[ix['Kernel Name']].split('(')[0][-28:], r[ix['Grid Size']]),{})[r[ix['Metric Name']]]=(r[ix['Metric Value']],r[ix['Metric Unit']])
for k,v in list(rec.items())[:8]: print(k, v)
PY
done
