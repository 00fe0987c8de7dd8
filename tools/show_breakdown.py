import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/breakdown.json'))
tot = sum(r['ms'] for r in d['rows'])
print("total ms", round(tot, 3))
agg = {}
for r in d['rows']:
    n = r['name']
    if n.startswith('b'):
        i = int(n[1:].split('.')[0])
        n = 'layer1' if i < 3 else 'layer2' if i < 7 else 'layer3' if i < 13 else 'layer4'
    a = agg.setdefault(n, [0, 0]); a[0] += r['ms']; a[1] += r['flops']
for k, (ms, fl) in agg.items():
    print(f"{k:8s} {ms:7.3f} ms {100*ms/tot:5.1f}%  {fl/(ms*1e-3)/1e12 if fl else 0:7.1f} TF/s")
