import sys, json
sys.path.insert(0, ".")
from tools.bench_configs import flat
from sfmlocalization_b200.gpu import HuloGpu
with HuloGpu(0) as g:
    for nB in (1250000, 2500000, 5000000, 10000000):
        flat(g, "4096 x %d" % nB, 4096, nB, 3000, steps=40)
