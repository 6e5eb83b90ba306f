"""Copies the evidence of the last `tools/gpu_final.sh` visit from gpurun_out/ into profiles/ (round-2 names) and
re-derives the ncu summaries.  Run in the build container after the visit."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
for src, dst in (('launches.csv', 'r2_launches.csv'), ('bench.log', 'r2_bench.json'), ('bench_ref.log', 'r2_bench_reference_arm.json'),
                 ('phase.log', 'r2_phase_profile.txt'), ('quick.log', 'r2_quick_bench.txt'), ('scheduler_bench.log', 'r2_scheduler_bench.json'),
                 ('sanitizer_racecheck.log', 'r2_sanitizer_racecheck.log'), ('sanitizer_synccheck.log', 'r2_sanitizer_synccheck.log'),
                 ('sanitizer_memcheck.log', 'r2_sanitizer_memcheck.log'), ('pytest_gpu.log', 'r2_pytest_gpu.log')):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
rep = os.path.join(G, 'prof.ncu-rep')
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
idx = [(k, hdr.index(k)) for k in keep if k in hdr]
with open(os.path.join(P, 'r2_ncu_full_summary.csv'), 'w') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit'] + [r[hdr.index('Kernel Name')][:40] for r in rows[2:]])
    for k, i in idx:
        w.writerow([k, units[i]] + [r[i] for r in rows[2:]])


def val(kname, metric):
    for r in rows[2:]:
        if kname in r[hdr.index('Kernel Name')]:
            return float(r[hdr.index(metric)]) * {'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'Gbyte': 1e9}[units[hdr.index(metric)]]


commit = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True, cwd=ROOT).stdout.strip()
tr = {k: int(val(k, 'dram__bytes_read.sum') + val(k, 'dram__bytes_write.sum')) for k in ('filter_kernel', 'sweep_kernel')}
tr.update(captured_at=f'commit {commit} (round 2, final kernels), ncu --set full, BASELINE config 4, one launch each',
          source='profiles/r2_ncu_full_summary.csv (dram__bytes_read.sum + dram__bytes_write.sum)')
json.dump(tr, open(os.path.join(P, 'traffic.json'), 'w'), indent=1)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
open('/tmp/src.csv', 'w').write(src)
a = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), '/tmp/src.csv', '40'], capture_output=True, text=True).stdout
b = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), '/tmp/src.csv', '40', 'samples'], capture_output=True, text=True).stdout
with open(os.path.join(P, 'r2_hot_lines.txt'), 'w') as f:
    f.write('# ncu --set full, config 4, one launch each of filter_kernel and sweep_kernel (both CTAs); CUDA source lines by executed warp instructions\n')
    f.write('\n'.join(l[:180] for l in a.splitlines()) + '\n\n# the same, by warp stall samples\n' + '\n'.join(l[:180] for l in b.splitlines()) + '\n')
print(open(os.path.join(P, 'r2_ncu_full_summary.csv')).read())
print(tr)
