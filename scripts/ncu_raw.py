"""Print the headline metrics of an `ncu --page raw --csv` dump."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
for r in rows[2:]:
    for w in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers']:
        if w in hdr: print(f"{w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
    for i,h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h:
            try: v=float(r[i].replace(',',''))
            except: continue
            if v>0.2: print('   stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),round(v,2))
