#!/usr/bin/env bash
# round 2, GPU call AP (1 GPU): per-kernel durations of the VCM passes of bench.py --config c5
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_vcm|k_film|k_light" -s 40 -c 24 --csv --log-file gpurun_out/c5_kernels.csv python bench.py --config c5 --steps 4 --warmup 2 > gpurun_out/c5_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/c5_kernels.csv') if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); m=h.index("Metric Name"); v=h.index("Metric Value"); u=h.index("Metric Unit"); i=h.index("ID")
per=collections.OrderedDict()
for r in rows[1:]:
    key=(r[i], r[k].split('(')[0]); per.setdefault(key,{})[r[m]]=(float(r[v].replace(',','')), r[u])
agg=collections.OrderedDict()
for (id_,name),d in per.items():
    t,unit=d["gpu__time_duration.sum"]; us=t*1e-3 if unit in('ns','nsecond') else (t if unit in ('us','usecond') else t*1e3)
    a=agg.setdefault(name,{"n":0,"us":0.0,"d":d}); a["n"]+=1; a["us"]+=us
for name,a in agg.items():
    d=a["d"]
    print("%-22s %2d launches %9.1f us/launch  regs %3.0f warps %4.1f%% lanes %4.1f issue %4.1f%% L1hit %4.1f%% dram r/w %.2f/%.2f GB" % (name,a["n"],a["us"]/a["n"],d["launch__registers_per_thread"][0],d["sm__warps_active.avg.pct_of_peak_sustained_active"][0],d["smsp__thread_inst_executed_per_inst_executed.ratio"][0],d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],d["l1tex__t_sector_hit_rate.pct"][0],d["dram__bytes_read.sum"][0]/(1e9 if d["dram__bytes_read.sum"][1]=='byte' else 1),d["dram__bytes_write.sum"][0]/(1e9 if d["dram__bytes_write.sum"][1]=='byte' else 1)))
PY
