"""Summary of one `ncu --set full --import-source on` capture of lbs_fuseg_kernel (read in the build container with `ncu -i`):
what the kernel moved through DRAM / the L2 crossbar / the SM's L1-shared data pipe, and where its warps waited.
  python tools/ncu_fuseg_summary.py gpurun_out/r03g_fuseg35_set_full.ncu-rep > profiles/r03g_fuseg35_summary.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]


def page(name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = page('raw')
names, units, vals = raw[0], raw[1], raw[-1]
m = {n: (v, u) for n, u, v in zip(names, units, vals)}


def num(n):
    return float(m[n][0].replace(',', ''))


want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'derived__lts__lts2xbar_bytes.sum.per_second',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']
print(f'# {rep}')
for n in want:
    if n in m:
        print(f'{n:80s} {m[n][0]:>18s} {m[n][1]}')
try:
    cyc = num('sm__cycles_elapsed.max') * 148
    tc = num('l1tex__data_pipe_tc_wavefronts_mem_shared.sum')
    lsu = num('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed')
    print(f'\nL1/shared data pipe, share of SM cycles: LSU (shared + global) {lsu:.1f} %, tensor-core operand reads {100 * tc / cyc:.1f} %'
          ' (+ the TMA writes of the operand ring: operand bytes per tile / 128 B per cycle)')
except Exception as e:      # older captures lack a metric
    print('(data-pipe shares unavailable:', e, ')')

src = page('source')
hdr = src[1]
ix = {n: i for i, n in enumerate(hdr)}
rows = src[2:]


def f(r, n):
    try:
        return float(r[ix[n]])
    except (ValueError, KeyError):
        return 0.0


tot = sum(f(r, '# Samples') for r in rows)
base = int(rows[0][ix['Address']], 16)
cols = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_mio', 'stall_lg', 'stall_sleep', 'stall_math', 'stall_not_selected',
        'stall_selected', 'stall_dispatch', 'stall_branch_resolving']
print(f'\nwarp-state samples: {int(tot)}; by reason: ' + ', '.join(f'{c[6:]} {100 * sum(f(r, c) for r in rows) / tot:.1f} %' for c in cols))
print('\ntop instructions by samples (offset, share, long/short scoreboard, executions, SASS):')
for r in sorted(rows, key=lambda r: -f(r, '# Samples'))[:16]:
    print(f"  {int(r[ix['Address']], 16) - base:#07x} {100 * f(r, '# Samples') / tot:5.1f} %  lsb {int(f(r, 'stall_long_sb')):5d} ssb {int(f(r, 'stall_short_sb')):5d}"
          f"  exec {int(f(r, 'Instructions Executed')):8d}  {r[ix['Source']][:70]}")
print('\nshared-memory wavefronts per instruction (executions, wavefronts, excess over the conflict-free count):')
for r in rows:
    w = f(r, 'L1 Wavefronts Shared')
    if w > 1e5:
        print(f"  {int(r[ix['Address']], 16) - base:#07x} exec {int(f(r, 'Instructions Executed')):8d} wavefronts {int(w):9d} excess {int(f(r, 'L1 Wavefronts Shared Excessive')):8d}  {r[ix['Source']][:56]}")
print('\nglobal-memory tag requests per instruction:')
for r in rows:
    t = f(r, 'L1 Tag Requests Global')
    if t > 1e5:
        print(f"  {int(r[ix['Address']], 16) - base:#07x} exec {int(f(r, 'Instructions Executed')):8d} requests {int(t):9d}  {r[ix['Source']][:56]}")
