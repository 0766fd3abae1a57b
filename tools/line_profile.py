"""Per-CUDA-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump:
warp instructions executed and stall samples per line, heaviest first."""
import csv
import sys


def main(path, top=45):
	rows = list(csv.reader(open(path)))
	fname = ''
	hdr = None
	lines = []
	for r in rows:
		if not r:
			continue
		if r[0] == 'File Path':
			fname = r[1].split('/')[-1]
			continue
		if r[0] == 'Line No':
			hdr = r
			iex = hdr.index('Instructions Executed')
			isamp = hdr.index('Warp Stall Sampling (All Samples)')
			continue
		if hdr is None or len(r) < len(hdr) or r[0] in ('', 'Function Name'):
			continue
		try:
			ln = int(r[0])
		except ValueError:
			continue
		num = lambda x: float(x) if x not in ('', '-', None) else 0.0
		# source text may contain quotes / commas that split the field: index the metrics from the right
		lines.append((fname, ln, r[1].strip(), num(r[iex - len(hdr)]), num(r[isamp - len(hdr)])))
	te = sum(l[3] for l in lines)
	ts = sum(l[4] for l in lines)
	print('total warp instructions %.1fM, samples %d' % (te / 1e6, ts))
	for l in sorted(lines, key=lambda l: -l[3])[:top]:
		print('%5.1f%% exec %5.1f%% stall  %s:%d  %s' % (100 * l[3] / te, 100 * l[4] / max(ts, 1), l[0], l[1], l[2][:100]))


if __name__ == '__main__':
	main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
