"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`): one line per launch with time, DRAM bytes, registers,
DRAM / issue utilisation and L2 hit rate.  `python tools/ncu_summary.py X.ncu-rep [out.json]`"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hd, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hd)}
SCALE = {'Gbyte': 1, 'Mbyte': 1e-3, 'Kbyte': 1e-6, 'byte': 1e-9, 'Tbyte': 1e3, 'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}
def val(r, k, scaled=False):
    try:
        v = float(r[idx[k]].replace(',', ''))
    except (ValueError, KeyError):
        return float('nan')
    return v * SCALE.get(units[idx[k]], 1) if scaled else v
print("%-46s %12s %8s %8s %8s %5s %6s %6s %6s %6s %8s" % ("kernel", "grid", "ms", "rd GB", "wr GB", "regs", "dram%", "L2hit%", "issue%", "warps%", "Ginst"))
out = []
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    short = name.split('(')[0].replace('void ', '').replace('opesci::', '')
    tp = name[name.find('<'):name.find('>') + 1].replace('(int)', '').replace('(bool)', '') if '<' in name else ''
    ms = val(r, 'gpu__time_duration.sum', True)
    rd, wr = val(r, 'dram__bytes_read.sum', True), val(r, 'dram__bytes_write.sum', True)
    rec = dict(kernel=short + tp, grid=r[idx['Grid Size']].replace(' ', ''), ms=ms, dram_read_GB=rd, dram_write_GB=wr,
               regs=int(val(r, 'launch__registers_per_thread')), dram_pct=val(r, 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed'),
               l2_hit_pct=val(r, 'lts__t_sector_hit_rate.pct'), issue_pct=val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
               warps_pct=val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'), ginst=val(r, 'smsp__inst_executed.sum') / 1e9)
    out.append(rec)
    print("%-46s %12s %8.3f %8.3f %8.3f %5d %6.1f %6.1f %6.1f %6.1f %8.3f" % (rec['kernel'][:46], rec['grid'], ms, rd, wr, rec['regs'], rec['dram_pct'],
                                                                        rec['l2_hit_pct'], rec['issue_pct'], rec['warps_pct'], rec['ginst']))
print("total: %.3f ms, %.2f GB read, %.2f GB written" % (sum(o['ms'] for o in out), sum(o['dram_read_GB'] for o in out), sum(o['dram_write_GB'] for o in out)))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
