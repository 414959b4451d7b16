"""Tabulate a window sweep (quick_bench jsonl) and the per-launch ncu metrics of the banded kernels."""
import sys, json, csv, collections
tag = sys.argv[1]
for l in open('gpurun_out/%s_random.jsonl' % tag):
    d = json.loads(l)
    print(d['band_env'], d['band_windows'], d['band_in_use'], 'primal %.3f dual %.3f iter %.3f it/s %.1f algoGB/s %.0f actual %.0f' % (
        d['primal_ms'], d['dual_ms'], d['ms_per_iter'], d['it_per_s'], d['algo_GBs'], d['actual_GBs']), d['band_ms'], d['setup_s'], d['device_GB'], d.get('band_shape'), d.get('band_shape_ms'))
try:
    rows = [r for r in csv.reader(l for l in open('gpurun_out/%s_random_ncu.csv' % tag) if l.startswith('"'))]
except FileNotFoundError:
    sys.exit(0)
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[idx['ID']], r[idx['Kernel Name']].split('(')[0][-25:])
    per.setdefault(key, {})[r[idx['Metric Name']]] = r[idx['Metric Value']]
def f(v, k, s=1e6): return float(v.get(k, 'nan')) / s
for k, v in per.items():
    print(k, 'us %.0f' % f(v, 'gpu__time_duration.sum', 1e3), 'req_tex %.1fM' % f(v, 'lts__t_requests_srcunit_tex.sum'),
          'sect_tex_rd %.1fM' % f(v, 'lts__t_sectors_srcunit_tex_op_read.sum'), 'fabric %.1fM' % f(v, 'lts__t_sectors_srcunit_ltcfabric.sum'),
          'l1req_active %s' % v.get('l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed'),
          'lts %s' % v.get('lts__throughput.avg.pct_of_peak_sustained_elapsed'), 'warps %s' % v.get('sm__warps_active.avg.pct_of_peak_sustained_active'),
          'l2hit %s' % v.get('lts__t_sector_hit_rate.pct'), 'dramR %.2fGB' % f(v, 'dram__bytes_read.sum', 1e9),
          'ls %s' % v.get('smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct'))
