#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-launch summary committed under profiles/.

  ncu -i gpurun_out/foo.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/foo.csv
"""
import csv
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max']


def main():
    rows = list(csv.reader(sys.stdin))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[i] for i in idx])
    w.writerow([units[i] for i in idx])
    for r in data:
        w.writerow([r[i] for i in idx])


if __name__ == '__main__':
    main()
