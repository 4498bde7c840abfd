"""Text summary of an .ncu-rep for profiles/: selected raw metrics per captured launch plus the hottest source
lines (stall samples / executed instructions).   python tools/ncu_summary.py report.ncu-rep [n_lines] > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu', '-i'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    head_ = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
    print(f'# {rep}  (ncu --set full --clock-control none --import-source on; repository HEAD when summarised: {head_})')
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}
    print('Kernel Name'.ljust(90), ' | '.join(r[col['Kernel Name']][:60] for r in data))
    for k in KEYS:
        if k in col:
            print(k.ljust(90), units[col[k]].ljust(16), ' | '.join(r[col[k]] for r in data))
    src = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass']))))
    out = []
    for r in src:
        if len(r) > 10 and r[0].isdigit():
            try:
                out.append((int(r[6]), int(r[7]), int(r[0]), r[1].strip()[:130]))
            except ValueError:
                pass
    ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print(f'\n# hottest source lines (first captured launch): % of stall samples, % of executed warp-instructions, line, source')
    seen = set()
    for o in sorted(out, reverse=True):
        if (o[2], o[3]) in seen:
            continue
        seen.add((o[2], o[3]))
        print('%5.1f%% %5.1f%%  %4d  %s' % (100 * o[0] / ts, 100 * o[1] / ti, o[2], o[3]))
        if len(seen) >= nlines:
            break
    # SASS opcode histogram of the captured kernel, weighted by executed warp-instructions (first captured launch)
    sass = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv', '--print-source', 'sass']))))
    hist, static, first = {}, {}, True
    for r in sass:
        if len(r) > 6 and r[0].startswith('0x'):
            toks = r[1].split()
            op = next((t for t in toks if not t.startswith('@')), '?').split('.')[0]
            try:
                ex = int(r[5])
            except ValueError:
                ex = 0
            hist[op] = hist.get(op, 0) + ex
            static[op] = static.get(op, 0) + 1
        elif r and r[0] == 'Kernel Name':
            if not first:
                break
            first = False
    tot = sum(hist.values()) or 1
    print(f'\n# SASS opcodes of the captured kernel: executed warp-instructions (% of {tot}), static count  [{sum(static.values())} instructions in the binary]')
    for op, ex in sorted(hist.items(), key=lambda kv: -kv[1])[:22]:
        print('%-10s %6.2f%%  %5d' % (op, 100.0 * ex / tot, static[op]))


if __name__ == '__main__':
    main()
