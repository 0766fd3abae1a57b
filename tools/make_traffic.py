"""profiles/traffic.json (DRAM bytes per launch of the hot kernels) from an `ncu --page raw --csv` dump of one
`ncu --set full` capture; bench.py reads it for roofline.traffic.  usage: make_traffic.py <hot_raw.csv> <source label>"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
	v = float(v)
	return int(round(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]))


def main(path, label):
	rows = list(csv.reader(open(path)))
	hdr, units = rows[0], rows[1]
	idx = {h: i for i, h in enumerate(hdr)}
	out = {}
	for r in rows[2:]:
		name = r[idx['Kernel Name']]
		key = 'k_pairs' if 'k_pairs' in name else ('k_rows2' if 'k_rows2' in name else None)
		if key is None or key in out:
			continue
		rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
		wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
		out[key] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr}
		g = lambda m: float(r[idx[m]]) if m in idx and r[idx[m]] not in ('', 'n/a') else None
		out[key]['limiter'] = {'l1tex_data_pipe_lsu_wavefronts_pct': g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),
			'issue_active_pct': g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
			'dram_throughput_pct': g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}
	sys.path.insert(0, ROOT)
	import bench
	json.dump({'source': label, 'kernel_source_sha256': bench.kernel_source_hash(), 'kernels': out}, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
	print(json.dumps(out))


if __name__ == '__main__':
	main(sys.argv[1], sys.argv[2])
