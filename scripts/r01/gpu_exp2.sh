#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for lim in 0 32 128; do
  echo "=== limit $lim (plain run)"; ./scripts/exp/fetch_granularity $lim
  echo "=== limit $lim (ncu)"
  ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file gpurun_out/fetch_$lim.csv ./scripts/exp/fetch_granularity $lim > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/fetch_$lim.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=0
d={}
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    d.setdefault((r[idi],r[ki]),{})[r[mi]]=float(r[vi].replace(',',''))
loads=148*1024*512
for (i,k),m in d.items():
    print("%-60s dramB/load %7.1f  l2sect/load %5.2f  l2req/load %5.2f  %8.3f ms"%(k[:60], m['dram__bytes_read.sum']*(1e9 if m['dram__bytes_read.sum']<1e6 else 1)/loads if False else 0, m['lts__t_sectors_srcunit_tex_op_read.sum']/loads, m['lts__t_requests_srcunit_tex_op_read.sum']/loads, m['gpu__time_duration.sum']), m['dram__bytes_read.sum'])
PY
done
