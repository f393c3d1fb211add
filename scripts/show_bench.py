"""Pretty-print the summary + JSON line produced by scripts/gpu_tests.sh."""
import json
import sys

for l in sys.stdin:
    l = l.strip()
    if l.startswith('{'):
        j = json.loads(l)
        print('train %.0f seq/s  %.4f ms/step | e2e %.0f | eval %.0f users/s (e2e %.0f) | eager %.3f ms | launches/step %.0f' % (
            j['value'], j['ms_per_step'], j['e2e']['value'], j['eval']['value'], j['eval']['e2e']['value'],
            j['eager_ms_per_step'], j['gpu_launches'] / j['steps']))
        for k, v in j['kernels'].items():
            print('  %-32s calls %5.1f  ms %.4f  share %.3f  gbs %s | in graph: %s us/launch  share %s  gbs %s' % (
                k, v['calls_per_step'], v['ms_per_step'], v['share'], v['gbs'], v.get('graph_us_per_launch'), v.get('graph_share'), v.get('graph_gbs')))
        print('  roofline', json.dumps(j.get('roofline')))
        if 'cpu_baseline' in j:
            print('  cpu_baseline', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
    elif l:
        print(l)
