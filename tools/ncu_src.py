"""Summarise the source page of one kernel of an .ncu-rep: per warp-role waits and top stall lines."""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', *(['--kernel-id', kid] if ':' in kid else ['--launch-skip', kid, '--launch-count', '1'])], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
S = lambda r: int(r[ix['# Samples']] or 0)
print(rows[0][:2], 'total samples', sum(S(r) for r in data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k, r in enumerate(data):
    s = r[ix['Source']]
    if any(t in s for t in ('UTCHMMA', 'UTCBAR', 'UBLKCP', 'LDTM', 'TRYWAIT')) and int(r[ix['Instructions Executed']] or 0) > 0:
        print(str(k).rjust(5), str(S(r)).rjust(6), r[ix['Instructions Executed']].rjust(9), s[:90])
print('--- top lines')
for k, r in sorted(enumerate(data), key=lambda kr: -S(kr[1]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(str(k).rjust(5), str(S(r)).rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']][:80], {h[6:]: r[ix[h]] for h in stall_cols if r[ix[h]] not in ('', '0')})
