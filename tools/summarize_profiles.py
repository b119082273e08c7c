"""Turns gpurun_out ncu artefacts into the committed summaries under profiles/.
   python tools/summarize_profiles.py <round-tag> <launches.csv | -> [<name>=<file.ncu-rep> ...]"""
import collections, csv, re, subprocess, sys, os

tag, launches = sys.argv[1], sys.argv[2]
reps = dict(a.split('=') for a in sys.argv[3:])
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles')
if launches != "-":
    rows = list(csv.reader(open(launches)))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('void ', '')
        agg[name][0] += 1
        agg[name][1] += float(r[ix['Metric Value']]) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out_dir, '%s_launches_summary.md' % tag), 'w') as f:
        f.write('# %s — every kernel launch of ONE 4K P49 vitl step (ncu `gpu__time_duration.sum`, `--clock-control none`)\n\n' % tag)
        f.write('Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv '
                '--log-file gpurun_out/launches.csv python tools/profile_step.py vitl m2 9` (eager launches, no CUDA '
                'graph; per-launch times are cold-cache and serialised: compare SHARES).\n\n')
        f.write('%d launches, %.1f ms summed kernel time.\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n' % (len(data), tot))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %d | %.2f | %.1f %% |\n' % (k, v[0], v[1], 100 * v[1] / tot))
    shutil_copy = os.path.join(out_dir, '%s_launches.csv' % tag)
    with open(shutil_copy, 'w') as f:
        w = csv.writer(f)
        w.writerow(['id', 'kernel', 'grid', 'block', 'ns'])
        for r in data:
            w.writerow([r[ix['ID']], re.sub(r'\(.*', '', r[ix['Kernel Name']]), r[ix['Grid Size']], r[ix['Block Size']], r[ix['Metric Value']]])
KEYS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.per_cycle_active',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'launch__shared_mem_per_block_dynamic']
# captures whose DRAM traffic bench.py reports as roofline.traffic: name -> (key, units of the captured launch)
TRAFFIC = {'halo_conv_up4_x9': ('pf_conv3_halo_kernel/up_conv_list.4.conv1', 9 * 392 * 518)}
import json
for name, rep in reps.items():
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    h, units, vals = rr[0], rr[1], rr[-1]
    with open(os.path.join(out_dir, '%s_%s_full.md' % (tag, name)), 'w') as f:
        f.write('# %s — `ncu --set full --clock-control none --import-source on` of `%s`\n\n' % (tag, vals[h.index('Kernel Name')]))
        f.write('source capture: %s\n\n| metric | value | unit |\n|---|---:|---|\n' % os.path.basename(rep))
        for i, m in enumerate(h):
            if any(m == k or m.startswith(k) for k in KEYS):
                f.write('| %s | %s | %s |\n' % (m, vals[i], units[i]))
    if name in TRAFFIC:
        key, n_units = TRAFFIC[name]

        mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
        tot_b = sum(float(vals[h.index(m)]) * mult[units[h.index(m)]] for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
        tp = os.path.join(out_dir, 'dram_traffic.json')
        d = json.load(open(tp)) if os.path.exists(tp) else {}
        d[key] = dict(bytes_per_unit=tot_b / n_units, unit='output pixel', launch_bytes=tot_b, launch_units=n_units,
                      source='profiles/%s_%s_full.md (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum)' % (tag, name))
        json.dump(d, open(tp, 'w'), indent=1)
    print('wrote', name)
