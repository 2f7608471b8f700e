"""Hot instructions of one launch in an .ncu-rep: python tools/ncu_hot.py rep [launch_index] [top]
Prints SASS lines sorted by stall samples with executed count and main stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
allrows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
rows = allrows[starts[li]:starts[li + 1]]
print(rows[0][1][:200])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot_s = sum(int(r[ix['# Samples']]) for r in body); tot_i = sum(int(r[ix['Instructions Executed']]) for r in body)
print('instructions', len(body), 'samples', tot_s, 'warp-inst executed', tot_i)
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print('stall totals', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:10])
if '--all' in sys.argv:
    for n, r in enumerate(body):
        print(n, r[ix['# Samples']], r[ix['Instructions Executed']], r[ix['Source']].strip())
    sys.exit()
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']]))[:top]
for i in sorted(order):
    r = body[i]
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls if int(r[ix[s]])), reverse=True)[:3]
    print(f"{i:5d} smp={r[ix['# Samples']]:>6} exe={r[ix['Instructions Executed']]:>8} {r[ix['Source']].strip()[:70]:70s} {st}")
