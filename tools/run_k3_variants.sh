# K3 at 4 / 5 / 6 / 8 resident pairs per SM (register cap 128 / 96 / 80 / 64): tools/bench_geom.py per variant
for mb in 4 5 6 8; do
  cp _variants/libhulo_mb$mb.so sfmlocalization_b200/libhulo_gpu.so   # built beforehand with -DHULO_GEO_MIN_BLOCKS=$mb into _variants/
  echo "== min blocks $mb"
  timeout 300 python tools/bench_geom.py 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['pairs'], d['matches_per_pair'], d['ransac_round'], d['outlier_frac'], d['device_ms'], d['valid'])
    else: print(l.rstrip())
"
done | tee gpurun_out/k3_variants.txt
