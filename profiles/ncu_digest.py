"""Digest of an ncu capture exported with `ncu -i X.ncu-rep --page raw --csv` / `--page source --csv`:
python profiles/ncu_digest.py X_raw.csv X_src.csv [block]  -> key counters, stall reasons, and a histogram of stall
samples / executed instructions over SASS address blocks (default 150 instructions)."""
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 150
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"{h} [{u}] = {v}")
        elif 'issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.05:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:22s} {float(v):.3f} per issue")
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if 'Source' in r and any('Sampling' in c for c in r):
        hdr, start = r, i + 1
        break
ix = {h: i for i, h in enumerate(hdr)}
data = rows[start:]
S = lambda a, b, key: sum(int(r[ix[key]] or 0) for r in data[a:b])
tot, ti = S(0, len(data), '# Samples'), S(0, len(data), 'Instructions Executed')
print(f"SASS instructions {len(data)}, stall samples {tot}, warp instructions executed {ti}")
for a in range(0, len(data), blk):
    b = min(a + blk, len(data))
    s_, i_ = S(a, b, '# Samples'), S(a, b, 'Instructions Executed')
    if s_ > 0.003 * tot or i_ > 0.003 * ti:
        print(f"{a:5d}-{b:5d} samples {100 * s_ / tot:5.1f}%  inst {100 * i_ / ti:5.1f}%  {data[a][ix['Source']].strip()[:60]}")
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("hot instructions (>0.8 % of the samples):")
for n, r in enumerate(data):
    s_ = int(r[ix['# Samples']] or 0)
    if s_ > tot * 0.008:
        st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
        print(f"{n:5d} {100 * s_ / tot:5.2f}% thr={r[ix['Avg. Threads Executed']][:5]:>5} {r[ix['Source']].strip()[:64]:64s} {st}")
