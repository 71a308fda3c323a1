"""Turn ncu captures (gpurun_out/*.ncu-rep, launches.csv) into the tracked summaries under profiles/.

    python tools/ncu_summary.py r01 gpurun_out/r01_launches.csv gpurun_out/r01_a.ncu-rep [more .ncu-rep ...]

Writes profiles/<tag>_ncu_launches_summary.md (per-kernel share of the step from the launch list),
profiles/<tag>_ncu_full_summary.md (key metrics of every --set full capture) and profiles/<tag>_traffic.json
(kernel -> DRAM bytes per launch, consumed by bench.py for roofline.traffic).
"""
import collections, csv, io, json, os, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def short(name):
    return name.split('(')[0].replace('void ', '').strip()


def main():
    args = sys.argv[1:]
    first = None
    if args[0] == "--first":
        first = int(args[1]); args = args[2:]
    tag, launches, reps = args[0], args[1], args[2:]
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles')
    # ---- launch list
    lines = [l for l in open(launches) if l.startswith('"')]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            if first is not None and n >= first:
                break
            n += 1
            agg.setdefault(short(row['Kernel Name']), []).append(float(row['Metric Value'].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list ({os.path.basename(launches)}): gpu__time_duration.sum per kernel\n\n",
           "Cold-cache, serialised launches under ncu: compare SHARES with bench.py's `kernels[].share`, not absolutes.\n",
           (f"Only the first {first} launches = the device-resident batch-512 steps (warm-up + timed); the rest of the CSV are the "
            "chunk launches of the host-buffer (e2e) pipeline, whose proportions differ.\n\n" if first else "\n"),
           "| kernel | launches | avg us | share of all profiled time |\n|---|---|---|---|\n"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |\n")
    open(os.path.join(root, f"{tag}_ncu_launches_summary.md"), 'w').write(''.join(out))
    # ---- full captures
    full = [f"# ncu --set full captures ({tag}); units are ncu's own\n"]
    traffic = {}
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(raw)))
        if len(r) < 3:
            continue
        hdr, units = r[0], r[1]
        for row in r[2:]:
            name = short(row[hdr.index('Kernel Name')])
            full.append(f"\n## {name}  ({os.path.basename(rep)})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    full.append(f"| {k} | {row[i]} | {units[i]} |\n")
            try:
                rd = float(row[hdr.index('dram__bytes_read.sum')].replace(',', '')) * UNIT[units[hdr.index('dram__bytes_read.sum')]]
                wr = float(row[hdr.index('dram__bytes_write.sum')].replace(',', '')) * UNIT[units[hdr.index('dram__bytes_write.sum')]]
                traffic[name] = rd + wr
            except Exception:
                pass
    open(os.path.join(root, f"{tag}_ncu_full_summary.md"), 'w').write(''.join(full))
    json.dump(traffic, open(os.path.join(root, f"{tag}_traffic.json"), 'w'), indent=1)
    print(''.join(out))
    print(json.dumps(traffic, indent=1))


if __name__ == '__main__':
    main()
