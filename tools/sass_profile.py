"""Summarise an `ncu --page source --csv` dump: where the executed instructions and stall samples are."""
import csv
import sys


def main(path, chunk=80):
	rows = list(csv.reader(open(path)))
	hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
	hdr = rows[hi]
	data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] != 'Address']
	isrc = hdr.index('Source')
	isamp = hdr.index('Warp Stall Sampling (All Samples)')
	iex = hdr.index('Instructions Executed')
	ithr = hdr.index('Avg. Threads Executed')
	num = lambda x: float(x) if x not in ('', None) else 0.0
	tot_s = sum(num(r[isamp]) for r in data)
	tot_e = sum(num(r[iex]) for r in data)
	print('SASS lines %d, samples %d, warp instructions executed %.1fM' % (len(data), tot_s, tot_e / 1e6))
	for k in range(0, len(data), chunk):
		ch = data[k:k + chunk]
		s = sum(num(r[isamp]) for r in ch)
		e = sum(num(r[iex]) for r in ch)
		thr = sum(num(r[ithr]) * num(r[iex]) for r in ch) / max(e, 1)
		ops = {}
		for r in ch:
			parts = r[isrc].split()
			if not parts:
				continue
			op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
			op = op.split('.')[0]
			ops[op] = ops.get(op, 0) + num(r[iex])
		top = sorted(ops.items(), key=lambda x: -x[1])[:6]
		print('%4d-%4d samples %5.1f%% exec %5.1f%% thr %4.1f  %s' % (k, k + chunk, 100 * s / max(tot_s, 1), 100 * e / max(tot_e, 1), thr,
			' '.join('%s:%.1fM' % (a, b / 1e6) for a, b in top)))
	print('top stalled instructions:')
	for r in sorted(data, key=lambda r: -num(r[isamp]))[:25]:
		print('  %5.2f%% exec %.2fM thr %4.1f  %s' % (100 * num(r[isamp]) / max(tot_s, 1), num(r[iex]) / 1e6, num(r[ithr]), r[isrc][:90]))


if __name__ == '__main__':
	main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 80)
