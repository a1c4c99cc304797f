"""Summarise an .ncu-rep (ncu --set full) into the few metrics we track.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [out.csv]"""
import csv, subprocess, sys
KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct', 'sm__warps_active.avg.pct', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'launch__waves', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__average_warps_issue_stalled',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__throughput.avg.pct',
        'lts__throughput.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__shared_mem_per_block')
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
import io
buf = io.StringIO(); wr = csv.writer(buf); wr.writerow(['kernel', 'metric', 'unit', 'value'])
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').replace('arfe::', '').replace('<unnamed>::', '')
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k) for k in KEEP):
            wr.writerow([name, h, u, v])
text = buf.getvalue().rstrip('\n')
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text + '\n')
else:
    print(text)
