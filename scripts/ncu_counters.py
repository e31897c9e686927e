"""Counter table of an ncu report (`ncu --set full --metrics ...`, see scripts/gpu/r03d.sh) and, optionally, the DRAM bytes per
launch of its kernels written into profiles/traffic.json (what bench.py reports as `roofline.traffic`).

usage: python scripts/ncu_counters.py <report.ncu-rep> <title> <particles per launch> [--traffic <workload> <source note>]
"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS = [('k_g2p2g', 'g2p2g'), ('k_g2p_adj', 'g2p_adj'), ('k_p2g_adj', 'p2g_adj'), ('k_grid_adj_flat', 'grid_op_adj'),
         ('k_grid_adj', 'grid_op_adj'), ('k_grid_flat', 'grid_op'), ('k_grid', 'grid_op'), ('k_p2g', 'p2g'),
         ('k_kinematics_adj', 'kinematics_adj'), ('k_kinematics', 'kinematics'), ('k_permute_rows', 'reorder')]
COLS = [('time ns', 'gpu__time_duration.sum'), ('warp instr', 'smsp__inst_executed.sum'), ('regs', 'launch__registers_per_thread'),
        ('issue active %', 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        ('warps active %', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('SM active cycles avg', 'sm__cycles_active.avg'), ('SM active cycles max', 'sm__cycles_active.max'),
        ('DRAM read B', 'dram__bytes_read.sum'), ('DRAM write B', 'dram__bytes_write.sum'),
        ('L2 RED sectors', 'lts__t_sectors_op_red.sum'), ('RED instr (warp)', 'smsp__inst_executed_op_global_red.sum'),
        ('smem atomics', 'smsp__inst_executed_op_shared_atom.sum'),
        ('smem bank conflicts', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
        ('smem wavefronts', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'),
        ('threads / instr', 'smsp__thread_inst_executed_per_inst_executed.ratio')]


def main():
    rep, title, particles = sys.argv[1], sys.argv[2], float(sys.argv[3])
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--print-units', 'base'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    groups = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r'\(SimConst.*', '', r[idx['Kernel Name']]).replace('void ', '')
        groups.setdefault(name, []).append(r)
    print(f'## {title}\n')
    print('| kernel | launches | ' + ' | '.join(c for c, _ in COLS) + ' | DRAM B / particle | thread-instr / particle |')
    print('|---|---|' + '---|' * (len(COLS) + 2))
    traffic = {}
    for name, rs in groups.items():
        def mean(metric):
            vals = [float(r[idx[metric]]) for r in rs if metric in idx and r[idx[metric]] not in ('', 'n/a')]
            return sum(vals) / len(vals) if vals else float('nan')
        vals = [mean(m) for _, m in COLS]
        dram = mean('dram__bytes_read.sum') + mean('dram__bytes_write.sum')
        print(f'| {name} | {len(rs)} | ' + ' | '.join(f'{v:.4g}' for v in vals) + f' | {dram / particles:.0f} | {mean("smsp__inst_executed.sum") * mean("smsp__thread_inst_executed_per_inst_executed.ratio") / particles:.0f} |')
        cls = next((c for p, c in CLASS if name.startswith(p)), None)
        if cls:
            traffic[cls] = (dram, name)
    if '--traffic' in sys.argv:
        i = sys.argv.index('--traffic')
        workload, note = sys.argv[i + 1], sys.argv[i + 2]
        path = os.path.join(ROOT, 'profiles', 'traffic.json')
        tr = json.load(open(path))
        for cls, (dram, name) in traffic.items():
            tr.setdefault(workload, {})[cls] = dict(
                dram_bytes_per_launch=dram,
                source=f'{note}: dram__bytes_read.sum + dram__bytes_write.sum of {name}, mean over the captured launches '
                       f'(ncu --set full --clock-control none; cold L2)')
        json.dump(tr, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()
