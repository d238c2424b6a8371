"""Key metrics + top stall sites of an ncu report (run where `ncu` is installed; no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main(path, ntop=14):
    raw = list(csv.reader(io.StringIO(run(['-i', path, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    for k, r in enumerate(raw[2:]):
        d = dict(zip(hdr, r))
        print('=== launch %d: %s grid %s block %s' % (k, d.get('Kernel Name', '')[:70], d.get('Grid Size'), d.get('Block Size')))
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                print('   %-82s %s %s' % (h, v, u))
    src = list(csv.reader(io.StringIO(run(['-i', path, '--page', 'source', '--csv']))))
    hi = [i for i, r in enumerate(src) if r and r[0] == 'Address']
    for k, h0 in enumerate(hi[:1]):
        h = src[h0]
        end = hi[k + 1] - 1 if k + 1 < len(hi) else len(src)
        data = [r for r in src[h0 + 1:end] if len(r) == len(h)]
        idx = {n: i for i, n in enumerate(h)}
        samp = idx['# Samples']
        stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
        tot = sum(int(r[samp] or 0) for r in data)
        agg = {c: sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
        print('--- stall samples (launch %d): total %d: %s' % (k, tot, sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
        for r in sorted(data, key=lambda r: -int(r[samp] or 0))[:ntop]:
            st = sorted([(int(r[idx[c]] or 0), c) for c in stall_cols], reverse=True)[:2]
            print('%7s  %-80s %s' % (r[samp], r[idx['Source']][:80], st))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
