"""Summarise an .ncu-rep (development aid): key metrics, stall mix, hottest SASS regions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
want = None
if "--kernel" in sys.argv:                       # pick the first launch whose name holds this substring (default: the first launch)
    k = sys.argv.index("--kernel"); want = sys.argv[k + 1]; del sys.argv[k:k + 2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(raw)) if r]
if len(rows) < 3:
    print("no kernel launches in", rep); sys.exit(0)
hdr, units = rows[0], rows[1]
iname = hdr.index('Kernel Name') if 'Kernel Name' in hdr else None
vals = next((r for r in rows[2:] if want is None or (iname is not None and want in r[iname])), rows[2])
if iname is not None: print("kernel:", vals[iname][:110])
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','smsp__inst_executed.sum',
 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct']
for i, h in enumerate(hdr):
    if h in keys: print("%-80s %-10s %s" % (h, units[i], vals[i]))
st = []
for i, h in enumerate(hdr):
    if 'smsp__pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
        try: st.append((float(vals[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError: pass
tot = sum(v for v, _ in st) or 1
print("stalls: " + "  ".join("%s %.1f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(src)) if r]
if len(rows) < 3 or 'Source' not in rows[1]:
    print("(no source page)"); sys.exit(0)
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, ins, ith, isam = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
tot = sum(int(r[ins]) for r in data); tots = sum(int(r[isam]) for r in data) or 1
regions, cur = [], None
for k, r in enumerate(data):
    n, s, t = int(r[ins]), int(r[isam]), float(r[ith] or 0)
    if cur and abs(cur['cnt'] - n) <= 0.15 * max(cur['cnt'], n, 1) + 1000:
        cur['n'] += 1; cur['tot'] += n; cur['samp'] += s; cur['end'] = k; cur['thr'] += t * n
    else:
        cur = {'start': k, 'end': k, 'n': 1, 'cnt': n, 'tot': n, 'samp': s, 'thr': t * n}; regions.append(cur)
print("total warp-instr", tot)
for g in regions:
    if g['tot'] > 0.006 * tot or g['samp'] > 0.01 * tots:
        print("sass %4d-%4d n=%3d each~%9d  instr %5.1f%%  samples %5.1f%%  thr %4.1f  %s" % (g['start'], g['end'], g['n'], g['cnt'], 100 * g['tot'] / tot, 100 * g['samp'] / tots, g['thr'] / max(g['tot'], 1), data[g['start']][isrc][:44]))
if len(sys.argv) > 2:  # dump a SASS range with per-instruction samples
    a, b = int(sys.argv[2]), int(sys.argv[3])
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    for k in range(a, b):
        r = data[k]
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("%4d %9s samp=%5s %-26s %s" % (k, r[ins], r[isam], ",".join("%s:%d" % (h, v) for v, h in top if v), r[isrc][:70]))
