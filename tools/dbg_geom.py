import sys, numpy as np
sys.path.insert(0, '.')
from oracle import oracle as orc
orc.build()
from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import HuloGpu
from tests.test_geom_gpu import batch
specs = [(16, 0.2), (17, 0.0), (24, 0.3), (40, 0.5), (64, 0.4), (100, 0.3), (129, 0.6), (200, 0.5),
         (333, 0.4), (512, 0.5), (700, 0.7), (1000, 0.6), (30, 1.0), (300, 1.0), (18, 0.1), (1500, 0.5)]
xI, xJ, off, truth = batch(specs, 100)
w, h = synth.IMAGE_WH
sizes = np.tile(np.array([w, h, w, h], np.int32), (len(specs), 1))
with HuloGpu(0) as g:
    for it in (5, 25, 200):
        r = g.geometric_filter(xI, xJ, off, sizes, 4.0, it, 7)
        print("max_iter", it)
        for p in range(len(specs)):
            a, b = int(off[p]), int(off[p + 1])
            o = orc.fmatrix_acransac(xI[a:b], xJ[a:b], (w, h), (w, h), 4.0, it, 7 + 1000003 * p)
            print(p, specs[p], bool(r['valid'][p]), o['ok'], len(r['inliers'][p]), len(o['inliers']),
                  round(float(r['nfa'][p]), 4), round(o['nfa'], 4),
                  np.array_equal(np.sort(r['inliers'][p]), np.sort(o['inliers'])),
                  round(float(r['error_max'][p]), 4), round(o['error_max'], 4))
