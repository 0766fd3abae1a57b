"""Key metrics + stall reasons from an `ncu --page raw --csv` dump."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
	'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
	'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
	'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__grid_size',
	'sm__cycles_elapsed.max', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
	'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'smsp__inst_executed_pipe_fp64.sum',
	'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum',
	'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
	'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
	'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main(path):
	rows = list(csv.reader(open(path)))
	hdr, units = rows[0], rows[1]
	idx = {h: i for i, h in enumerate(hdr)}
	st = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
	for r in rows[2:]:
		print('=====', r[idx['Kernel Name']][:60])
		for k in KEYS:
			if k in idx:
				print('  %-66s %s %s' % (k, r[idx[k]], units[idx[k]]))
		vals = sorted([(float(r[idx[h]] or 0), h) for h in st], reverse=True)[:7]
		print('  stalls: ' + ', '.join('%s %.2f' % (h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v) for v, h in vals))


if __name__ == '__main__':
	main(sys.argv[1])
