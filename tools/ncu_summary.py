"""Print the metrics that matter from an .ncu-rep (raw page) -- developer helper, run on the CPU box."""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
                       'sm__pipe_tensor', 'sm__inst_executed_pipe_uniform', 'sm__warps_active.avg.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
                       'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum ', 'lts__throughput',
                       'l1tex__data_bank_conflicts_pipe_lsu', 'l1tex__throughput', 'smsp__average_warp', 'smsp__issue_active.avg.pct',
                       'sm__throughput.avg.pct', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg ', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
                       'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'smsp__warp_issue_stalled', 'dram__cycles_active',
                       'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'sm__sass_inst_executed_op_shared', 'l1tex__m_xbar2l1tex_read_bytes', 'smsp__inst_executed.sum ']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for data in rows[2:]:
    print('==', data[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, data):
        if any(p.strip() in h for p in pats):
            print(f"  {h:75s} {u:12s} {v}")
