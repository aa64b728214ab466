#!/usr/bin/env python
"""Selected columns of an `ncu --page raw --csv` export (one line per profiled launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum']
idx = [(w, hdr.index(w) if w in hdr else -1) for w in want]
units = rows[1]
for r in rows[2:]:
    print(' | '.join('%s=%s%s' % (w.split('.')[0].split('__')[-1][:22], r[i][:46] if i >= 0 else 'NA', (' ' + units[i]) if i >= 0 and units[i] and w != 'Kernel Name' else '') for w, i in idx))
