"""Summarise an .ncu-rep: key raw metrics + top stall locations from the source page."""
import csv, subprocess, sys, io
rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_l1tex2xbar_write_bytes.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__cluster_size',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print("kernel:", r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
si, so = hdr.index('# Samples'), hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot)
agg = {}
for r in data:
    for i in stall:
        if int(r[i]) > 0:
            agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i])
print("stall reasons, whole kernel: " + ", ".join(f"{k} {100 * v / max(1, sum(agg.values())):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))
for idx, r in enumerate(data):
    r.append(idx)
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    st = {hdr[i][6:]: int(r[i]) for i in stall if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[si]):6d} {100 * int(r[si]) / tot:5.1f}%  #{r[-1]:4d} {r[so].strip()[:64]:64s} {st}")
