"""Per-section instruction breakdown of a profiled kernel: joins the per-SASS-instruction execution counts of an ncu
report (`--set full --import-source on`) with the source lines nvdisasm recovers from the library's -lineinfo, and
aggregates them by code section (file + line range table below).

usage: python scripts/ncu_sections.py <report.ncu-rep> <kernel regex> <library.so> <particles per launch> [launch index]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

# (file suffix, first line, last line, section); first match wins.  Line ranges refer to the sources of the commit the
# profiled library was built from -- the table is regenerated from function names, see section_of().
FUNC_SECTIONS = [
    ('svd3.cuh', 'jacobi_pair|swapneg|svd3\\b', 'svd (Jacobi sweeps, sort, U)'),
    ('svd3.cuh', 'svd3_backward|clamp_gap', 'svd backward'),
    ('particle_math.cuh', 'von_mises', 'return map (log/exp, U e V^T)'),
    ('particle_math.cuh', 'p2g_particle_impl|p2g_particle\\b|p2g_particle_adj', 'F_tmp, stress, affine'),
    ('particle_math.cuh', 'bspline1|make_stencil', 'stencil weights / offsets'),
    ('particle_math.cuh', 'g2p_particle|g2p_finish|g2p_gather', 'g2p gather (27 nodes)'),
    ('particle_math.cuh', 'p2g_adj_finish', 'p2g.grad finish (stress / return-map adjoints)'),
    ('particle_math.cuh', 'p2g_adj_particle|p2g_adj_gather', 'p2g.grad gather (27 nodes)'),
    ('particle_math.cuh', 'g2p_adj_begin|g2p_adj_node', 'g2p.grad scatter values'),
    ('particle_math.cuh', 'g2p_adj_finish', 'g2p.grad gather (weight adjoints)'),
    ('particle_math.cuh', 'load_m3|load_v3|store_m3|store_v3|store_svd', 'particle loads / stores'),
    ('kernels_common.cuh', 'warp_scatter|butterfly|slot_value|f4shfl|f4add|f4sel|red_add4|scatter27|scatter9', 'scatter reduction (shuffles / smem + RED)'),
    ('kernels_common.cuh', 'mark_tile|mark_stencil_tiles|node_offset', 'active-tile marking'),
    ('mpm_math.cuh', '.*', 'small 3x3 / vector helpers'),
]


def function_ranges(path):
    """line ranges of the DSK_DEV / __global__ functions of a source file (brace matching from the signature)."""
    src = open(path).read().split('\n')
    out = []
    i = 0
    sig = re.compile(r'^(?:template.*\n)?\s*(?:DSK_DEV|__global__|static|inline).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
    while i < len(src):
        m = re.match(r'^\s*(?:DSK_DEV|__global__)\b.*?([A-Za-z_][A-Za-z0-9_]*)\s*\(', src[i])
        if not m and src[i].strip().startswith('k_') and i > 0 and '__global__' in src[i - 1]:
            m = re.match(r'^\s*([A-Za-z_][A-Za-z0-9_]*)\s*\(', src[i])
        if m:
            name = m.group(1)
            depth, j, seen = 0, i, False
            while j < len(src):
                depth += src[j].count('{') - src[j].count('}')
                seen = seen or '{' in src[j]
                if seen and depth <= 0:
                    break
                j += 1
            out.append((i + 1, j + 1, name))
            i = j + 1
        else:
            i += 1
    return out


def main():
    rep, kern, lib, particles = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    launch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    csvtxt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--kernel-name',
                             f'regex:{kern}', '--launch-skip', str(launch), '--launch-count', '1'], capture_output=True, text=True).stdout
    lines = csvtxt.split('\n')
    kname = next(l for l in lines if l.startswith('"Kernel Name"'))
    mangled_hint = re.search(r'"void (\w+)', kname) or re.search(r'"(\w+)', kname.split(',', 1)[1])
    rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    hdr = rows[0]
    ia, ii, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    isrc = hdr.index('Source')
    stall_cols = [(h, hdr.index(h)) for h in ('# Samples', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_lg',
                                               'stall_mio', 'stall_math', 'stall_not_selected', 'stall_selected',
                                               'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch') if h in hdr]
    base = int(rows[1][ia], 16)
    counts = {}
    stalls = {}
    for r in rows[1:]:
        if len(r) <= it or not r[ia].startswith('0x'):
            continue
        counts[int(r[ia], 16) - base] = (int(r[ii]), int(r[it]), r[isrc].split()[0] if r[isrc].split() else '')
        stalls[int(r[ia], 16) - base] = [int(r[c] or 0) for _, c in stall_cols]
    # nvdisasm with line info
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # pick the function whose demangled name matches
    tpl = re.search(r'<\(int\)(\d+)(?:, \(bool\)(\d))?>', kname)
    fn = (re.search(r'void (\w+)', kname) or re.search(r'"Kernel Name","(\w+)', kname)).group(1)
    sec_re = re.compile(r'^\s*\.section\s+\.text\.(\S+?),')
    cur, cur_line, per_line = None, None, collections.Counter()
    stall_line = collections.defaultdict(lambda: [0] * len(stall_cols))
    want = None
    for l in dis.split('\n'):
        m = sec_re.match(l)
        if m:
            cur = m.group(1)
            ok = re.search(r'\d+' + fn + r'(I|\d|P|v)', cur) is not None
            if ok and tpl:
                ok = f'ILi{tpl.group(1)}E' in cur and (tpl.group(2) is None or f'Lb{tpl.group(2)}E' in cur)
            want = cur if ok and want is None else (want if want != cur else want)
            continue
        if cur != want or want is None:
            continue
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S+)', l)
        if m and cur_line:
            off = int(m.group(1), 16)
            if off in counts:
                per_line[cur_line] += counts[off][1]
                for q, v in enumerate(stalls[off]):
                    stall_line[cur_line][q] += v
    # sections by enclosing function
    csrc = os.environ.get('NCU_SRC_DIR') or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'diffskill_b200', 'csrc')
    ranges = {f: function_ranges(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith(('.cuh', '.cu'))}

    def section_of(f, ln):
        name = next((n for a, b, n in ranges.get(f, []) if a <= ln <= b), '?')
        for suffix, pat, sec in FUNC_SECTIONS:
            if f == suffix and re.fullmatch(pat, name):
                return sec
        return f'{f}:{name}'

    sec = collections.Counter()
    for (f, ln), c in per_line.items():
        sec[section_of(f, ln)] += c
    total = sum(sec.values())
    tot_all = sum(c[1] for c in counts.values())
    print(f'# {kname.strip()[:140]}')
    print(f'# thread-instructions per particle: {tot_all / particles:.0f} (mapped to source lines: {total / particles:.0f})')
    print('| section | thread-instr / particle | share |')
    print('|---|---|---|')
    for s, c in sec.most_common():
        print(f'| {s} | {c / particles:.0f} | {100.0 * c / total:.1f} % |')
    if stall_cols:
        # warp-state samples of the same launch by section (the sampler's view of where the warps WAIT)
        ssec = collections.defaultdict(lambda: [0] * len(stall_cols))
        for (f, ln), v in stall_line.items():
            sc = section_of(f, ln)
            for q, x in enumerate(v):
                ssec[sc][q] += x
        tot = sum(v[0] for v in ssec.values()) or 1
        print('\n| section | ' + ' | '.join(('samples %' if h == '# Samples' else h.replace('stall_', '')) for h, _ in stall_cols) + ' |')
        print('|---|' + '---|' * len(stall_cols))
        for sc, v in sorted(ssec.items(), key=lambda kv: -kv[1][0]):
            print(f'| {sc} | {100.0 * v[0] / tot:.1f} | ' + ' | '.join(f'{100.0 * x / tot:.1f}' for x in v[1:]) + ' |')
        if os.environ.get('NCU_TOP_LINES'):
            print('\n| file:line | samples % | long_sb % |')
            print('|---|---|---|')
            for (f, ln), v in sorted(stall_line.items(), key=lambda kv: -kv[1][0])[:int(os.environ['NCU_TOP_LINES'])]:
                print(f'| {f}:{ln} | {100.0 * v[0] / tot:.1f} | {100.0 * v[1] / tot:.1f} |')
    ops = collections.Counter()
    for off, (wi, ti, op) in counts.items():
        ops[op.split('.')[0]] += ti
    print('\n| opcode | thread-instr / particle |')
    print('|---|---|')
    for o, c in ops.most_common(18):
        print(f'| {o} | {c / particles:.0f} |')


if __name__ == '__main__':
    main()
